"""Multi-GPU correctness of the sharded fit on real devices (runs when the box has >= 2 GPUs; the world-size-2 gloo
tests in test_dist_cpu.py cover the host logic everywhere).  One process per GPU under torchrun, NCCL:
tools/dist_fit_check.py asserts that the row-sharded cMLEimat (fused predict -> Gram, ONE all-reduce of [G|C|zz],
replicated K x K stage) equals the single-GPU fit of the same data to 1e-10, that the NVLink peer-memory sum equals
NCCL's all-reduce of the same buffers to 1e-13 and is bitwise identical on every rank."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fit_equals_single_gpu_fit(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs on this box (gpurun --gpus {world})")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), str(ROOT / "tools" / "dist_fit_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"dist_fit_check_n{world}.json").write_text(json.dumps(res, indent=1))
    assert res["world"] == world and res["ok"], res
    assert res["dM"] < 1e-10 and res["dw"] < 1e-10 and res["dV"] < 1e-10 and res["dnll"] < 1e-12, res
    assert res["peer_allreduce_ok"] is True and res["peer_vs_nccl_packed_max_rel_diff"] < 1e-13, res
