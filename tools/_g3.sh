cd $GRAFT_REPO_ROOT
python -m pytest tests/test_predict_gpu.py tests/test_ref_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/k1_smallk_probe.py gpurun_out/k1_probe_r02.json
python bench.py --steps 3 --warmup 3 --no-fit > gpurun_out/bench_nofit.json 2> gpurun_out/bench_nofit.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mrts_predict_kernel -s 3 -c 1 -o gpurun_out/k1_r02_bench -f \
   python bench.py --steps 1 --warmup 3 --no-fit > gpurun_out/k1_r02_bench_ncu.log 2>&1; echo "ncu rc=$?"
