cd $GRAFT_REPO_ROOT
python -m pytest tests/test_frk_gpu.py tests/test_large_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/gram_probe.py > gpurun_out/gram_probe_r02_final.json 2> gpurun_out/gram_probe.err; tail -3 gpurun_out/gram_probe.err
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tma_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/gram_r02_wide -f python tools/gram_ncu.py 2000000,500,100 > gpurun_out/gram_r02_wide_ncu.log 2>&1; echo "ncu rc=$?"
