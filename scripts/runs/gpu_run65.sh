mkdir -p gpurun_out/r2y
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2y/pytest_final.log 2>&1; tail -3 gpurun_out/r2y/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y/smoke_final.log 2>&1; tail -1 gpurun_out/r2y/smoke_final.log
