cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu > gpurun_out/r2_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t14.log
tail -4 gpurun_out/r2_t14.log
python tools/e2e_probe.py 4000000 > gpurun_out/e2e3.log 2>&1; tail -2 gpurun_out/e2e3.log | cut -c1-1200
