cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t9.log
tail -4 gpurun_out/r2_t9.log
FRK_HOST_STAGING=0 python tools/e2e_probe.py 1000000 4000000 > gpurun_out/e2e0.log 2>&1; tail -3 gpurun_out/e2e0.log | cut -c1-330
python tools/e2e_probe.py 1000000 4000000 > gpurun_out/e2e1.log 2>&1; tail -3 gpurun_out/e2e1.log | cut -c1-330
FRK_HOST_COPY_THREADS=8 python tools/e2e_probe.py 4000000 > gpurun_out/e2e2.log 2>&1; tail -2 gpurun_out/e2e2.log | cut -c1-330
