# N-GPU sequence on the GPU box: `gpurun --gpus N -- "bash tools/_g2.sh N"` (sharded-fit tests, then the bench under torchrun)
cd $GRAFT_REPO_ROOT
N=$1
python -m pytest tests/test_dist_gpu.py tests/test_large_gpu.py -q -m gpu > gpurun_out/r2_dist_n$N.log 2>&1; tail -4 gpurun_out/r2_dist_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
b=json.load(open("gpurun_out/bench_r02_n$N.json"))
print("value",b["value"],"ms",b["ms_per_step"],"e2e",b["e2e"]["value"],"pageable",b["e2e"]["pageable"]["value"])
print("fused",{k:v["value"] for k,v in b["e2e"]["fused"].items() if isinstance(v,dict)})
for c in b["fit"]["cmle_sharded"]: print({k:c[k] for k in ("T","predict_gram_ms","collective_ms","nccl_allreduce_ms","peer_sum_ms","kxk_stage_ms","wall_ms","collective","check","collective_fallback_reason")})
PY
