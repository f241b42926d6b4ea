"""One Gram launch per shape for the profiler (C4 shape and the narrow K=32 shape).  Measurement tooling."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from autofrk_b200 import _lib

shapes = ((2_000_000, 500, 100), (5_000_000, 32, 4))
if len(sys.argv) > 1:  # n,K,T n,K,T ...
    shapes = tuple(tuple(int(v) for v in arg.split(",")) for arg in sys.argv[1:])
for n, K, T in shapes:
    F = torch.randn((K, n), dtype=torch.float64, device="cuda").T
    Z = torch.randn((T, n), dtype=torch.float64, device="cuda").T
    packed = torch.empty(_lib.lib.frk_gram_packed_len(K, T), dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(int(os.environ.get("REPS", "2"))):
        _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F.data_ptr()), n, K, n, C.c_void_p(Z.data_ptr()), T, n, None,
                                               C.c_void_p(packed.data_ptr()), st))
    torch.cuda.synchronize()
    del F, Z
print("ok")
