"""Run under torchrun on N GPUs: row-sharded cMLEimat (fused predict->Gram, one all-reduce, replicated K x K
stage) must equal the single-GPU fit of the same data.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from autofrk_b200 import dist as D
from autofrk_b200 import mrts as G


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    d, m, K, T, n = 2, 400, 80, 6, 300_001
    rng = np.random.default_rng(0)
    Xu = rng.random((m, d))
    fit = G.mrtsrcpp(Xu, G.xobs_diag_of(Xu, m), K - d - 1) if rank == 0 else None
    fit = D.broadcast_fit(fit, src=0)                       # bit-identical eigenvectors on every rank
    basis = G.MrtsBasis(fit["X"], fit["UZ"], Xu, fit["nconst"], fit["BBBH"])
    plan = basis.plan()
    g = torch.Generator(device="cuda").manual_seed(5)       # same global data on every rank, then sliced
    x_all = torch.rand((d, n), dtype=torch.float64, device="cuda", generator=g)
    z_all = torch.randn((T, n), dtype=torch.float64, device="cuda", generator=g)
    r0, r1 = D.shard_rows(n, world, rank)
    x, Z = x_all[:, r0:r1].contiguous().T, z_all[:, r0:r1].contiguous().T
    out = D.fit_sharded(plan, x, Z, n_total=n, s=0.0, wSave=True)          # the collective as one NCCL all-reduce
    res = {}
    peer_ok, peer_err, packed_diff = None, "", None
    try:   # the same collective as our own kernel over NVLink peer memory
        from autofrk_b200 import _lib
        peer = D.PeerAllReduce(_lib.lib.frk_gram_packed_len(K, T), x.device)
        # the packed statistics themselves: peer-memory sum against NCCL's all-reduce of the same per-rank buffers
        D.predict_gram_shard(plan, x, Z, out=peer.buf)
        nccl_sum = peer.buf.clone()
        dist.all_reduce(nccl_sum, op=dist.ReduceOp.SUM)
        peer_sum = peer().clone()
        packed_diff = float((peer_sum - nccl_sum).abs().max() / nccl_sum.abs().max())
        ref_sum = peer_sum.clone()
        dist.broadcast(ref_sum, src=0)
        packed_bitwise = torch.tensor([int(torch.equal(ref_sum, peer_sum))], device="cuda")
        dist.all_reduce(packed_bitwise, op=dist.ReduceOp.MIN)
        outp = D.fit_sharded(plan, x, Z, n_total=n, s=0.0, wSave=True, peer=peer)
        outp2 = D.fit_sharded(plan, x, Z, n_total=n, s=0.0, wSave=True, peer=peer)      # buffer reuse
        relp = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        peer_ok = bool(relp(outp["M"], out["M"]) < 1e-12 and relp(outp["w"], out["w"]) < 1e-12
                       and np.array_equal(outp2["M"], outp["M"]))
        # every rank must hold bitwise-identical sums
        chk = torch.tensor(np.ascontiguousarray(outp["M"])).cuda()
        ref0 = chk.clone()
        dist.broadcast(ref0, src=0)
        peer_ok = peer_ok and bool(torch.equal(chk, ref0)) and bool(packed_bitwise.item() == 1) and packed_diff < 1e-13
    except Exception as e:  # pragma: no cover
        peer_err = repr(e)[:300]
    if rank == 0:
        # single-GPU reference of the same call (world-size-independent path)
        dist_initialized = True
        packed = D.predict_gram_shard(plan, x_all.T, z_all.T)
        torch.cuda.synchronize()
        from autofrk_b200.frk import cmle_from_stats
        ref = cmle_from_stats(packed[:K * K], packed[K * K:K * K + K * T], float(packed[-1].item()), n, T, s=0.0,
                              wSave=True, onlylogLike=False, K=K, on_device=True)
        rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        res = dict(world=world, v=[out["v"], ref["v"]], dM=rel(out["M"], ref["M"]), dw=rel(out["w"], ref["w"]),
                   dV=rel(out["V"], ref["V"]), dnll=abs(out["negloglik"] - ref["negloglik"]) / abs(ref["negloglik"]))
        res["ok"] = bool(res["dM"] < 1e-10 and res["dw"] < 1e-10 and res["dV"] < 1e-10 and res["dnll"] < 1e-12)
        res["collective_nccl_vs_single_gpu"] = "ok" if res["ok"] else "MISMATCH"
        res["peer_allreduce_ok"] = peer_ok
        res["peer_vs_nccl_packed_max_rel_diff"] = packed_diff
        res["peer_err"] = peer_err
        res["ok"] = bool(res["ok"] and peer_ok is not False)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not res["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
