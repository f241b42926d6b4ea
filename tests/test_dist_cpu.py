"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, the packed-statistics
all-reduce and the fit broadcast.  The device kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n, K, T, q):
    import torch
    import torch.distributed as dist

    import oracle as O
    from autofrk_b200 import dist as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        rng = np.random.default_rng(0)                       # identical data on both ranks
        F = rng.standard_normal((n, K))
        Z = rng.standard_normal((n, T))
        r0, r1 = D.shard_rows(n, world_size, rank)
        G, Cm, zz = O.gram_stats(F[r0:r1], Z[r0:r1])          # stand-in for frk_gram_stats_dev on the shard
        packed = torch.from_numpy(np.concatenate([G.ravel(order="F"), Cm.ravel(order="F"), [zz]]))
        D.allreduce_stats(packed)
        Gf, Cf, zzf = O.gram_stats(F, Z)
        Gu, Cu, zu = D.unpack_stats(packed.numpy(), K, T)
        ok = (np.allclose(Gu, Gf, rtol=1e-13, atol=1e-10) and np.allclose(Cu, Cf, rtol=1e-13, atol=1e-10)
              and abs(float(zu) - zzf) < 1e-9 * zzf)
        fit = {"X": rng.standard_normal((7, 5)), "UZ": rng.standard_normal((10, 5)),
               "BBBH": rng.standard_normal((3, 7)), "nconst": rng.random(2)} if rank == 0 else None
        got = D.broadcast_fit(fit, src=0)
        ref = np.random.default_rng(0)
        ref.standard_normal((n, K)); ref.standard_normal((n, T))
        ok = ok and np.array_equal(got["X"], ref.standard_normal((7, 5))) and got["UZ"].shape == (10, 5) \
            and got["BBBH"].shape == (3, 7) and got["nconst"].shape == (2,)
        q.put((rank, bool(ok), (r0, r1)))
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition():
    from autofrk_b200.dist import shard_rows

    for n in (1, 2, 7, 100, 12_500_001, 100_000_000):
        for ws in (1, 2, 4, 8):
            blocks = [shard_rows(n, ws, r) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
            assert all(b[0] % 2 == 0 or b[1] == b[0] for b in blocks)     # non-empty shards start on an even row
            assert sum(b[1] - b[0] for b in blocks) == n


def test_allreduce_and_broadcast_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 2001, 9, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    blocks = dict((r, b) for r, _, b in res)
    assert blocks[0] == (0, 1002) and blocks[1] == (1002, 2001)
