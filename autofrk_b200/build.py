"""In-tree build of libfrk_b200.so (nvcc, sm_100a only).  (The checker builds itself: oracle/build_oracle.py.)

`python -m autofrk_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
Objects are rebuilt only when a source or header is newer (so repeated calls are cheap).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libfrk_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # stream 0 in this library = the calling host thread's own default stream: the K x K stages (launches without a
    # stream argument, blocking copies) of two host threads do not serialise on the legacy null stream
    "--default-stream", "per-thread",
    "-Xcompiler", "-fPIC",
    # tuning aid: extra flags for an experimental build, e.g. FRK_NVCC_EXTRA="-DFRK_EXPERIMENTAL_CONFIGS" 
    *os.environ.get("FRK_NVCC_EXTRA", "").split(),
]
LINK_LIBS = ["-lcusolver", "-lcublas"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(map(str, cmd[:3])))
    return r.stdout + r.stderr


def build_cuda(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "frk_b200.h", Path(__file__)]
    jobs = []
    objs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)] + (["-Xptxas", "-v"] if verbose else []))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_run, jobs):
                if verbose:
                    print(out)
    if force or jobs or _stale(LIB, objs):
        _run([NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xlinker", "-rpath,/usr/local/cuda/lib64", *LINK_LIBS])
    return LIB


def main():
    verbose = "-v" in sys.argv
    force = "-f" in sys.argv
    print(build_cuda(verbose=verbose, force=force))


if __name__ == "__main__":
    main()
