cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu > gpurun_out/r2_t12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t12.log
python tools/gram_probe.py > gpurun_out/r2_gram_probe.log 2>&1
cp gpurun_out/gram_probe.json gpurun_out/gram_probe_r02_v9.json
tail -4 gpurun_out/r2_t12.log; grep '"path": "default"' gpurun_out/r2_gram_probe.log | cut -c1-130
