cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_r02_n1.json')); print(b['value'], b['roofline']['frac'], b['roofline']['traffic'], b['e2e']['value'], b['e2e']['pageable']['value'], {k:(v['value'],v['wall_ms'],v['wall_ms_mean']) for k,v in b['e2e']['fused'].items() if isinstance(v,dict)})"
