"""Quick K4 timing probe (device-resident F, Z).  Measurement tooling."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from autofrk_b200 import _lib


def run(n, K, T, reps=3, path=""):
    import os
    os.environ.pop("FRK_GRAM_PATH", None)
    if path:
        os.environ["FRK_GRAM_PATH"] = path
    F = torch.randn((K, n), dtype=torch.float64, device="cuda").T
    Z = torch.randn((max(T, 1), n), dtype=torch.float64, device="cuda").T
    plen = _lib.lib.frk_gram_packed_len(K, T)
    packed = torch.empty(plen, dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def go():
        _lib.check(_lib.lib.frk_gram_stats_dev(C.c_void_p(F.data_ptr()), n, K, n, C.c_void_p(Z.data_ptr()), T, n, None,
                                               C.c_void_p(packed.data_ptr()), st))
    go()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = n * K * (K + 1.0) + 2.0 * n * K * T + 2.0 * n * T
    res = dict(n=n, K=K, T=T, path=path or "default", ms=best, tflops_credited=flops / best / 1e9, gbs=8.0 * n * (K + T) / best / 1e6)
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "small":      # only the shapes of the register-resident kernel (A/B of its tuning aids)
        out = [run(5_000_000, K, T, reps=5, path="mid")
               for K, T in ((24, 1), (32, 4), (48, 8), (64, 4), (64, 1), (80, 1), (100, 1), (128, 4), (128, 16), (100, 20))]
        json.dump(out, open("gpurun_out/gram_probe_small.json", "w"), indent=1)
        sys.exit(0)
    out = [run(5_000_000, 500, 100), run(2_000_000, 500, 1), run(1_000_000, 1000, 100),
           run(5_000_000, 200, 20), run(5_000_000, 256, 1), run(2_000_000, 446, 1)]
    # shapes default autoFRK() selects (K = 18 ... ~130): the register-resident kernel against the ones it replaces
    for K, T in ((8, 1), (16, 1), (24, 1), (32, 4), (48, 8), (64, 4), (64, 1), (100, 1), (128, 4), (128, 16), (100, 20)):
        out.append(run(5_000_000, K, T))
        out.append(run(5_000_000, K, T, path="mid"))
        if K <= 48 and T <= 16:
            out.append(run(5_000_000, K, T, path="narrow"))
        out.append(run(5_000_000, K, T, path="tma"))
    json.dump(out, open("gpurun_out/gram_probe.json", "w"), indent=1)
