mkdir -p gpurun_out/r2w
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2w/pytest.log 2>&1; tail -3 gpurun_out/r2w/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w/smoke.log 2>&1; tail -1 gpurun_out/r2w/smoke.log
python bench.py > gpurun_out/r2w/bench.json 2> gpurun_out/r2w/bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w/bench_ref.json 2> gpurun_out/r2w/bench_ref.err; echo ref rc=$?
bash scripts/ncu_r02.sh list > gpurun_out/r2w/list.log 2>&1
