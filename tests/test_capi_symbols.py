"""CPU-only: the C-ABI library loads and exports every symbol include/frk_b200.h declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "frk_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frk_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from autofrk_b200 import build

    lib = ctypes.CDLL(str(build.LIB))
    names = declared_symbols()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    from autofrk_b200 import _lib

    assert sorted(_lib._SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_device():
    """Without a GPU every compute entry point must fail loudly (FRK_ECUDA), never compute on the CPU."""
    import numpy as np
    import pytest
    from autofrk_b200 import _lib
    from autofrk_b200 import mrts as G

    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    Xu = np.random.default_rng(0).random((10, 2))
    with pytest.raises(_lib.FrkError) as ei:
        G.mrtsrcpp_predict(Xu, np.eye(2), Xu, np.zeros((3, 10)), np.zeros((13, 5)), np.ones(2), 2)
    assert ei.value.code == _lib.FRK_ECUDA


def test_product_never_imports_oracle():
    for p in (ROOT / "autofrk_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_header_is_plain_c():
    """include/frk_b200.h must parse as C (no C++ in the ABI).  (The R `.Call` shim is checked in test_r_shim.py.)"""
    import subprocess

    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-x", "c", str(ROOT / "include" / "frk_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "error" not in r.stderr


def test_gram_job_plan_covers_every_block_pair_once():
    """Host-side packing of the Gram work (no device): every 32-column block pair (a <= b) of F'F and every
    F x Z pair appears in exactly one job; a job holds <= 16 pairs over <= 8 distinct blocks."""
    import numpy as np
    from autofrk_b200 import _lib

    lib = _lib.lib
    for K, T in [(49, 0), (64, 4), (100, 1), (200, 20), (256, 1), (446, 1), (500, 1), (500, 100), (1000, 100), (1500, 33)]:
        njobs = ctypes.c_int32(0)
        count = lib.frk_gram_job_plan(K, T, ctypes.byref(njobs), None, 0)
        tri = np.zeros((count, 3), dtype=np.int32)
        assert lib.frk_gram_job_plan(K, T, ctypes.byref(njobs), tri.ctypes.data, count) == count
        kb, zb = -(-K // 32), -(-T // 32)
        want = {(a, b) for a in range(kb) for b in range(a, kb)} | {(a, kb + z) for a in range(kb) for z in range(zb)}
        got = [(int(a), int(b)) for _, a, b in tri]
        assert len(got) == len(set(got)) and set(got) == want, (K, T)
        assert tri[:, 0].max() == njobs.value - 1
        for j in range(njobs.value):
            sel = tri[tri[:, 0] == j]
            assert 1 <= len(sel) <= 16, (K, T, j)
            assert len(set(sel[:, 1]) | set(sel[:, 2])) <= 8, (K, T, j)
        # never more work than the 128-column tiling it replaced (16 warp tiles per tile pair, 10 on the diagonal)
        ntf, ntz = -(-K // 128), -(-T // 128)
        assert count <= 10 * ntf + 16 * (ntf * (ntf - 1) // 2 + ntf * ntz)


def test_gram_mid_plan_covers_every_tile_once_and_balances_the_pipes():
    """Host-side plan of the register-resident Gram kernel (no device): every 8 x 8 tile of the upper triangle of F'F
    and of F'Z is held by exactly one (warp, accumulator slot); the four SM sub-partitions issue the same number of
    DMMAs (within one per k4-step); shapes that do not fit are declined."""
    import numpy as np
    from autofrk_b200 import _lib

    lib = _lib.lib
    lib.frk_gram_mid_plan.restype = ctypes.c_int64
    for K in list(range(1, 129)) + [129, 200, 256]:
        for T in (0, 1, 4, 16, 20, 32, 33, 100):
            ns = ctypes.c_int32(0)
            sp = (ctypes.c_int32 * 4)()
            cnt = lib.frk_gram_mid_plan(K, T, ctypes.byref(ns), sp, None, 0)
            MT, NZ = -(-K // 8), -(-T // 8)
            kb, zb = -(-K // 32), -(-T // 32)
            if cnt < 0:
                assert K > 128 or T > 32, (K, T)      # every K <= 128 with T <= 32 fits: 32 units, 8 column blocks
                continue
            tiles = np.zeros((cnt, 5), dtype=np.int32)
            assert lib.frk_gram_mid_plan(K, T, ctypes.byref(ns), sp, tiles.ctypes.data, cnt) == cnt
            want = {(a, b, 0) for a in range(MT) for b in range(a, MT)} | {(a, z, 1) for a in range(MT) for z in range(NZ)}
            got = [(int(a), int(b), int(z)) for a, b, z, _, _ in tiles]
            assert len(got) == len(set(got)) and set(got) == want, (K, T)
            holders = [(int(w), int(s)) for _, _, _, w, s in tiles]
            assert len(set(holders)) == cnt and all(0 <= w < 16 // ns.value and 0 <= s < 16 for w, s in holders), (K, T)
            assert ns.value in (1, 2, 4) and sum(sp) == cnt * ns.value
            assert max(sp) - min(sp) <= 2, (K, T, list(sp))


def test_k1_traffic_record_was_captured_from_the_committed_kernel_sources():
    """bench.py reports roofline.traffic from profiles/k1_traffic_r02.json only while the SHA-256 of the K1 sources
    equals the one recorded at capture time (tools/k1_traffic_capture.py).  The committed record must match the
    committed sources -- otherwise the driver-run line would carry traffic = null."""
    import json
    import sys

    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    import bench

    rec = json.loads((root / "profiles" / "k1_traffic_r02.json").read_text())
    assert rec["k1_sources"] == list(bench.K1_SOURCES)
    assert rec["k1_sources_sha256"] == bench.sources_sha256(bench.K1_SOURCES)
    assert 0.9 < rec["traffic_bytes"] / rec["algorithmic_bytes"] < 1.1


def test_library_is_sm100a_code_with_fp64_tensor_and_tma_instructions():
    """The shipped binary is what DESIGN.md says it is: sm_100a cubins only, FP64 tensor-core instructions
    (DMMA.8x8x4; tcgen05/TMEM have no f64 kind), TMA bulk and tensor copies (UBLKCP, UTMALDG) and mbarrier traffic."""
    import shutil
    import subprocess

    import pytest

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    lib = Path(__file__).resolve().parent.parent / "autofrk_b200" / "libfrk_b200.so"
    sass = subprocess.run([cuobjdump, "-sass", str(lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    assert sass.count("DMMA.8x8x4") > 1000
    assert "UBLKCP.S.G" in sass and "UTMALDG.2D" in sass and "SYNCS.ARRIVE.TRANS64" in sass
