# one-GPU round-end sequence on the GPU box (run through gpurun: `gpurun -- 'bash tools/_g1.sh'`); outputs in gpurun_out/
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('gpurun_out/bench_r02_n1.json')); print(b['value'], b['roofline']['frac'], b['roofline']['traffic'], b['e2e']['value'], b['e2e']['pageable']['value'], {k:(v['value'],v['wall_ms'],v['wall_ms_mean']) for k,v in b['e2e']['fused'].items() if isinstance(v,dict)})"
