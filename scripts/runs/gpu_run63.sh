mkdir -p gpurun_out/r2y
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2y/pytest.log 2>&1; tail -3 gpurun_out/r2y/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y/smoke.log 2>&1; tail -2 gpurun_out/r2y/smoke.log
python bench.py > gpurun_out/r2y/bench.json 2> gpurun_out/r2y/bench.err; echo bench rc=$?
