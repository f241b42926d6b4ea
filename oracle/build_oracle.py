"""Builds the checker: the plain-C restatement (libfrk_oracle.so) and, where the upstream sources are present,
oracle/_ref/libautofrk_ref.so = the reference's own src/autoFRK.cpp compiled UNMODIFIED from where it lies
behind the stand-in headers of oracle/stub/.  TEST INFRASTRUCTURE: building the checker is not using it.

    python -m oracle.build_oracle          (or: make -C oracle all ref)

The reference sources are never copied into this repository; only the built library (git-ignored) travels to
the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ODIR = Path(__file__).resolve().parent
REF_SRC = Path(os.environ.get("FRK_REFERENCE_ROOT", "/root/reference")) / "src" / "autoFRK.cpp"
ORACLE_LIB = ODIR / "libfrk_oracle.so"
REF_LIB = ODIR / "_ref" / "libautofrk_ref.so"


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).exists() and Path(d).stat().st_mtime > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("oracle build step failed")


def build_oracle() -> Path:
    """gcc -O2 (R's default optimisation level) on the scalar restatement."""
    src = ODIR / "tps_scalar.c"
    if _stale(ORACLE_LIB, [src]):
        _run([os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-o", str(ORACLE_LIB), str(src), "-lm"])
    return ORACLE_LIB


def build_ref():
    """g++ -O2 on oracle/ref_wrap.cpp, which #includes the upstream translation unit.  Returns the library path,
    or None when the upstream sources are absent (GPU box) and no prebuilt library travelled."""
    deps = [ODIR / "ref_wrap.cpp", *sorted((ODIR / "stub").glob("*.h")), REF_SRC]
    if REF_SRC.exists():
        if _stale(REF_LIB, deps):
            REF_LIB.parent.mkdir(exist_ok=True)
            _run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{ODIR / 'stub'}",
                  f'-DFRK_REF_SRC="{REF_SRC}"', "-o", str(REF_LIB), str(ODIR / "ref_wrap.cpp")])
        return REF_LIB
    return REF_LIB if REF_LIB.exists() else None


if __name__ == "__main__":
    print(build_oracle())
    print(build_ref())
