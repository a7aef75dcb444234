mkdir -p gpurun_out/r2o
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2o/pytest.log 2>&1; tail -3 gpurun_out/r2o/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o/smoke.log 2>&1; tail -2 gpurun_out/r2o/smoke.log
python bench.py > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench.err; echo bench rc=$?
