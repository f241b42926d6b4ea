"""ctypes binding of libfrk_b200.so (the C ABI in include/frk_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import time,
and every compute entry point fails with FRK_ECUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FRK_B200_LIB", _PKG / "libfrk_b200.so"))

FRK_OK, FRK_EINVAL, FRK_EINTERNAL, FRK_ECUDA, FRK_ENOTCONV = 0, 1, 2, 3, 4
FRK_OUT_X1, FRK_OUT_BASIS = 0, 1


class FrkError(RuntimeError):
    """Raised for every non-zero status of the C ABI; `.code` holds the FRK_E* value."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA library first (python -m autofrk_b200.build). "
        "autofrk_b200 has no CPU fallback."
    )

lib = C.CDLL(str(LIB_PATH))

_dp = C.POINTER(C.c_double)
_i64, _i32, _vp = C.c_int64, C.c_int32, C.c_void_p

_SIGNATURES = {
    "frk_last_error": (C.c_char_p, []),
    "frk_version": (_i32, []),
    "frk_device_count": (_i32, [C.POINTER(_i32)]),
    "frk_set_device": (_i32, [_i32]),
    "frk_release_cached_memory": (_i32, []),
    "frk_measure_dmma_peak": (_i32, [_i32, C.POINTER(C.c_double)]),
    "frk_mrtsrcpp_predict": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _i32, _vp, _i64, _i64, _vp, _i64, _i64,
                                    _vp, _i64, _i32, _vp]),
    "frk_mrtsrcpp": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "frk_mrtsrcpp_predict0": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp,
                                     _vp]),
    "frk_getASCeigens": (_i32, [_vp, _i32, _vp, _vp]),
    "frk_plan_create": (_i32, [C.POINTER(_vp), _vp, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _i32, _i32]),
    "frk_plan_destroy": (_i32, [_vp]),
    "frk_plan_info": (_i32, [_vp, C.POINTER(_i64)]),
    "frk_plan_launches": (_i64, [_vp]),
    "frk_plan_predict": (_i32, [_vp, _vp, _i64, _vp, _i32]),
    "frk_plan_predict_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _i32, _vp]),
    "frk_gram_stats": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "frk_gram_packed_len": (_i64, [_i32, _i32]),
    "frk_gram_job_plan": (_i64, [_i32, _i32, _vp, _vp, _i64]),
    "frk_gram_mid_plan": (_i64, [_i32, _i32, _vp, _vp, _vp, _i64]),
    "frk_gram_stats_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _i64, _vp, _vp, _vp]),
    "frk_plan_predict_gram_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, _i32, _i64, _vp, _vp, _i64, _vp]),
    "frk_peer_sum_dev": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "frk_getHalf": (_i32, [_vp, _i64, _i32, _vp]),
    "frk_cMLEimat_stats": (_i32, [_vp, _i64, _vp, _i64, C.c_double, _i32, _i32, _i64, C.c_double, _i32, _i32, _i32,
                                  _vp, _vp, _vp, _vp, _vp]),
    "frk_cMLEimat_sweep": (_i32, [_vp, _i64, _vp, _i64, C.c_double, _i32, _i64, C.c_double, _vp, _i32, _vp]),
    "frk_cMLE": (_i32, [_vp, _i64, _i32, _i32, C.c_double, _vp, _vp, C.c_double, C.c_double, _i32, _i32, _i32,
                        C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "frk_plan_predict_frk": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "frk_invCz": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "frk_getLikelihood": (_i32, [_vp, _vp, _vp, _vp, C.c_double, _vp, _i64, _i32, _i32, _vp]),
    "frk_krige_from_stats": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    "frk_plan_predict_gram": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "frk_masked_stats": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "frk_EM0miss_stats": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, C.c_double, _i32, C.c_double, _i32, C.c_double,
                                 _vp, _vp, _vp, _vp]),
    "frk_knn_impute": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32]),
    "frk_basis_predict_frk": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
}


def _bind():
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


_bind()


def check(status: int):
    if status != FRK_OK:
        raise FrkError(status, lib.frk_last_error().decode("utf-8", "replace"))


def fmat(a, name="matrix") -> np.ndarray:
    """float64, column-major (R layout), 2-D; copies only when needed."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError(f"{name} must be a matrix")
    return np.asfortranarray(a)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def device_count() -> int:
    n = _i32(0)
    check(lib.frk_device_count(C.byref(n)))
    return n.value
