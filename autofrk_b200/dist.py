"""Row-sharded multi-GPU driver: one process per GPU (torchrun), torch.distributed for the plumbing.

The hot path shards by rows (locations are independent): knots / eigenvectors / Cpoly are replicated,
basis evaluation needs no collective at all, and the fit needs exactly one -- the sum of the packed
sufficient statistics [G (K*K) | C (K*T) | zz] (2.4 MB at K=500, T=100), after which the K x K stage runs
replicated on every rank.  The sum runs as our own kernel over NVLink peer memory (PeerAllReduce ->
frk_peer_sum_dev: symmetric-memory buffers, P2P loads, rank-ordered addition, bitwise-identical on all
ranks); NCCL's all-reduce (allreduce_stats) is the fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def bind_to_gpu_numa(device_index: int) -> int:
    """Pin this process to the CPUs NVML reports as local to the GPU, before any pinned host buffer is allocated,
    so that the host side of the H2D / D2H copies (predict.mrts returns an n2 x K host matrix) stays on the GPU's
    own socket.  Returns the number of CPUs bound to, or 0 when nothing was changed (no NVML, or the local CPUs
    are outside this process's cpuset)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            tok = visible.split(",")[device_index].strip()
            handle = pynvml.nvmlDeviceGetHandleByUUID(tok) if tok.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(tok))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(ncpu, 1024) + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus or cpus == allowed:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def shard_rows(n: int, world_size: int, rank: int):
    """Contiguous row block of `rank`; block starts are even so that every shard of a column-major
    matrix with even leading dimension stays 16-byte aligned (the Gram kernel's cp.async fast path)."""
    per = -(-n // world_size)
    per += per & 1
    r0 = min(n, rank * per)
    return r0, min(n, r0 + per)


def unpack_stats(packed, K: int, T: int):
    """views (G K x K, C K x T, zz) of a packed statistics buffer (torch tensor or numpy array)."""
    G = packed[:K * K].reshape(K, K).T if hasattr(packed, "is_cuda") else packed[:K * K].reshape(K, K, order="F")
    Cm = (packed[K * K:K * K + K * T].reshape(T, K).T if hasattr(packed, "is_cuda")
          else packed[K * K:K * K + K * T].reshape(K, T, order="F"))
    return G, Cm, packed[K * K + K * T]


def allreduce_stats(packed):
    """sum the packed statistics over all ranks (in place); no-op for a single process."""
    import torch.distributed as dist

    ws, _ = world()
    if ws > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed


def broadcast_fit(fit: dict | None, src: int = 0) -> dict:
    """Rank `src` fitted the knots (mrtsrcpp); everybody gets bit-identical X, UZ, BBBH, nconst."""
    import torch
    import torch.distributed as dist

    ws, rank = world()
    if ws == 1:
        return fit
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    keys = ("X", "UZ", "BBBH", "nconst")
    shapes = [None]
    if rank == src:
        shapes = [[list(np.shape(fit[k])) for k in keys]]
    dist.broadcast_object_list(shapes, src=src)
    out = {}
    for k, shp in zip(keys, shapes[0]):
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(fit[k], dtype=np.float64).T)).to(dev)
        else:
            t = torch.empty(list(reversed(shp)), dtype=torch.float64, device=dev)
        dist.broadcast(t, src=src)
        out[k] = np.asfortranarray(t.cpu().numpy().T)
    return out


class PeerAllReduce:
    """The fit's single collective as our own kernel over NVLink peer memory (frk_peer_sum_dev): the packed statistics
    live in a symmetric-memory allocation, every rank reads all peers' buffers with P2P loads and adds them in rank
    order.  torch's symmetric memory provides the allocation, the peer mappings and the device-side barriers."""

    def __init__(self, length: int, device):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.length = int(length)
        self.buf = symm.empty(self.length, dtype=torch.float64, device=device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.out = torch.empty(self.length, dtype=torch.float64, device=device)
        self.world = self.hdl.world_size
        self._ptrs = (C.c_void_p * self.world)(*[C.c_void_p(int(p)) for p in self.hdl.buffer_ptrs])

    def __call__(self):
        """`self.buf` holds this rank's statistics (written on the current stream); returns the sum over ranks."""
        import torch

        self.hdl.barrier(channel=0)                                # every rank's statistics are complete and visible
        st = C.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)
        check(lib.frk_peer_sum_dev(self._ptrs, self.world, self.length, C.c_void_p(self.out.data_ptr()), st))
        self.hdl.barrier(channel=1)                                # peers are done reading before the next overwrite
        return self.out


def predict_gram_shard(plan, x_t, Z_t, chunk_rows: int = 0, w_t=None, out=None):
    """packed statistics of this rank's rows, basis never materialised (frk_plan_predict_gram_dev)."""
    import torch

    n2, d = x_t.shape
    T = 0 if Z_t is None else Z_t.shape[1]
    K = plan.k
    packed = out if out is not None else torch.empty(lib.frk_gram_packed_len(K, T), dtype=torch.float64, device=x_t.device)
    ldx = x_t.stride(1) if d > 1 else n2
    ldz = (Z_t.stride(1) if T > 1 else n2) if T else n2
    st = C.c_void_p(torch.cuda.current_stream(x_t.device).cuda_stream)
    check(lib.frk_plan_predict_gram_dev(plan._h, C.c_void_p(x_t.data_ptr()), n2, ldx,
                                        C.c_void_p(Z_t.data_ptr()) if T else None, T, ldz,
                                        None if w_t is None else C.c_void_p(w_t.data_ptr()),
                                        C.c_void_p(packed.data_ptr()), int(chunk_rows), st))
    return packed


def fit_sharded(plan, x_t, Z_t, n_total: int, s: float = 0.0, wSave: bool = True, chunk_rows: int = 0, peer=None):
    """cMLEimat over row shards: local fused predict->Gram, ONE collective, replicated K x K stage.
    `peer` (a PeerAllReduce of the right length) runs the collective as our own NVLink peer-memory kernel;
    otherwise it is one NCCL all-reduce.  Returns the same list as cMLEimat (R/autoFRK.R:399-480)."""
    import torch

    from .frk import cmle_from_stats

    K, T = plan.k, Z_t.shape[1]
    if peer is not None:
        predict_gram_shard(plan, x_t, Z_t, chunk_rows, out=peer.buf)
        packed = peer()
    else:
        packed = allreduce_stats(predict_gram_shard(plan, x_t, Z_t, chunk_rows))
    torch.cuda.current_stream().synchronize()
    zz = float(packed[-1].item())
    G = packed[:K * K]
    Cm = packed[K * K:K * K + K * T]
    return cmle_from_stats(G, Cm, zz, n_total, T, s=s, wSave=wSave, onlylogLike=False, K=K, on_device=True)
