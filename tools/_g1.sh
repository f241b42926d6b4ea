cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_r02_n1.json')); print(b['value'], b['roofline']['frac'], b['roofline']['traffic'], b['e2e']['value'], b['e2e']['pageable']['value'], {k:v['value'] for k,v in b['e2e']['fused'].items() if isinstance(v,dict)})"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_n1.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_r02_reference_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/bench_launches_r02.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
