"""Builds r/src/frk_b200_shim.c against the mock of R's C API (tests/r_stub/mock_r.c) and drives it from Python the
way `.Call` would (test infrastructure)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "tests" / "r_stub" / "libfrk_rshim_mock.so"
REALSXP, INTSXP, LGLSXP, VECSXP, NILSXP = 14, 13, 10, 19, 0


def build():
    srcs = [ROOT / "tests" / "r_stub" / "mock_r.c", ROOT / "r" / "src" / "frk_b200_shim.c"]
    deps = srcs + [ROOT / "tests" / "r_stub" / "Rinternals.h", ROOT / "include" / "frk_b200.h"]
    if not SO.exists() or any(p.stat().st_mtime > SO.stat().st_mtime for p in deps):
        lib_dir = ROOT / "autofrk_b200"
        r = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-fPIC", "-shared", f"-I{ROOT / 'tests' / 'r_stub'}", f"-I{ROOT / 'include'}",
                            "-o", str(SO), *map(str, srcs), f"-L{lib_dir}", "-lfrk_b200", f"-Wl,-rpath,{lib_dir}"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return SO


class MockR:
    def __init__(self):
        L = C.CDLL(str(build()))
        for f in ("mock_real_matrix", "mock_real_vector", "mock_int_vector", "mock_int_matrix", "mock_nil", "mock_list_elt"):
            getattr(L, f).restype = C.c_void_p
        L.mock_real_ptr.restype = C.POINTER(C.c_double)
        L.mock_length.restype = C.c_longlong
        for f in ("mock_routine_name", "mock_last_error", "mock_list_name"):
            getattr(L, f).restype = C.c_char_p
        for f in ("mock_type", "mock_length", "mock_nrow", "mock_ncol", "mock_real_ptr"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.mock_list_elt.argtypes = [C.c_void_p, C.c_int]
        L.mock_list_name.argtypes = [C.c_void_p, C.c_int]
        L.mock_real_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mock_real_vector.argtypes = [C.c_void_p, C.c_ssize_t]
        L.mock_int_vector.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int]
        L.mock_int_matrix.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.mock_load()
        self.L = L

    def routines(self):
        return {self.L.mock_routine_name(i).decode(): self.L.mock_routine_nargs(i) for i in range(self.L.mock_nroutines())}

    # R values -------------------------------------------------------------------------------------------------
    def mat(self, a):
        a = np.asfortranarray(np.asarray(a, dtype=np.float64))
        return C.c_void_p(self.L.mock_real_matrix(a.ctypes.data, a.shape[0], a.shape[1]))

    def vec(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
        return C.c_void_p(self.L.mock_real_vector(a.ctypes.data, a.size))

    def ivec(self, a, logical=False):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())
        return C.c_void_p(self.L.mock_int_vector(a.ctypes.data, a.size, int(logical)))

    def imat(self, a):
        a = np.asfortranarray(np.asarray(a, dtype=np.int32))
        return C.c_void_p(self.L.mock_int_matrix(a.ctypes.data, a.shape[0], a.shape[1]))

    def nil(self):
        return C.c_void_p(self.L.mock_nil())

    def to_py(self, s):
        L = self.L
        t = L.mock_type(s)
        if t == NILSXP:
            return None
        if t == REALSXP:
            n = L.mock_length(s)
            a = np.ctypeslib.as_array(L.mock_real_ptr(s), shape=(max(n, 1),))[:n].copy()
            nr, nc = L.mock_nrow(s), L.mock_ncol(s)
            return a.reshape((nr, nc), order="F") if nr >= 0 else a
        if t == VECSXP:
            return {L.mock_list_name(s, i).decode(): self.to_py(C.c_void_p(L.mock_list_elt(s, i))) for i in range(L.mock_length(s))}
        raise TypeError(f"unexpected SEXP type {t}")

    def call(self, name, *args):
        """.Call(name, ...): returns the result converted to numpy / dict; raises RuntimeError(message) on Rf_error."""
        arr = (C.c_void_p * len(args))(*args)
        out = C.c_void_p()
        if self.L.mock_dotcall(name.encode(), len(args), arr, C.byref(out)):
            raise RuntimeError(self.L.mock_last_error().decode())
        return self.to_py(out)

    def free(self):
        self.L.mock_free_all()
