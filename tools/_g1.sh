cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:mrts_predict_kernel -s 3 -c 1 -f -o gpurun_out/k1_r02_bench python bench.py --steps 1 --warmup 3 --no-fit > gpurun_out/k1_ncu.log 2>&1; echo "ncu rc=$?"
