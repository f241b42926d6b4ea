cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu > gpurun_out/r2_t16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t16.log; tail -3 gpurun_out/r2_t16.log
python tools/gram_probe.py > gpurun_out/r2_gram_probe.log 2>&1
cp gpurun_out/gram_probe.json gpurun_out/gram_probe_r02_v11.json
REPS=1 ncu --set full --import-source on --clock-control none -k regex:gram_mid_kernel -c 3 -f -o gpurun_out/gram_mid_r02_v11 python tools/gram_ncu.py 5000000,32,4 5000000,64,4 5000000,100,1 > gpurun_out/r2_ncu_mid.log 2>&1
grep '"path": "default"' gpurun_out/r2_gram_probe.log | cut -c1-125
