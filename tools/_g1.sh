cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu > gpurun_out/r2_t10.log 2>&1; tail -3 gpurun_out/r2_t10.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_n1.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
ncu --set full --clock-control none --import-source on -k regex:mrts_predict_kernel -s 3 -c 1 -f -o gpurun_out/k1_r02_bench python bench.py --steps 1 --warmup 3 --no-fit > gpurun_out/k1_ncu.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/bench_launches_r02.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu_list.log 2>&1; echo "ncu list rc=$?"
cut -c1-1500 gpurun_out/bench_r02_n1.json
