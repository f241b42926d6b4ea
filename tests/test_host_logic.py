"""CPU-only: the host mirror's argument handling (autofrk_b200/frk.py) -- the checks the reference makes, or the failures
it runs into, BEFORE any n-sized work (R/autoFRK.R:1168-1245), and the small host-side rules (K grid).  No device
call is reached: these run without a GPU."""
import numpy as np
import pytest

from autofrk_b200 import frk as G


def _obj(K=3, T=2, n=5):
    return {"w": np.zeros((K, T)), "V": np.eye(K), "M": np.eye(K), "s": 1.0, "G": np.ones((n, K))}


def test_predict_FRK_rejects_what_the_reference_rejects():
    with pytest.raises(ValueError, match='should use the option "wsave=TRUE"'):            # :1171-1173
        G.predict_FRK({"M": np.eye(2)}, newloc=np.zeros((3, 2)))
    with pytest.raises(ValueError, match="obsloc was given without obsData"):               # yhat undefined in R (:1214, :1248)
        G.predict_FRK(_obj(), obsloc=np.zeros((3, 2)))
    with pytest.raises(ValueError, match="Basis matrix of new locations should be given"):  # :1178-1180: G is not an mrts object
        G.predict_FRK(_obj(), newloc=np.zeros((3, 2)))
    with pytest.raises(ValueError, match="Dimensions of obsloc and obsData are not compatible"):   # :1199-1201
        G.predict_FRK(_obj(), obsData=np.zeros(4))


def test_predict_FRK_missing_data_model_fails_where_r_fails():
    """R/autoFRK.R:1224-1245: miss[, tt] runs past the pattern's columns when the fit dropped empty replicates (object$w keeps
    their columns, :834-836); solve(object$s * De) is singular at s = 0 (:847).  Both are raised before any device work."""
    m = _obj(T=2)
    m["missing"] = {"miss": np.zeros((5, 1))}
    m["pinfo"] = {"pick": np.arange(5), "D": np.ones(5)}
    with pytest.raises(IndexError, match="subscript out of bounds"):
        G.predict_FRK(m, se_report=True)
    m["w"] = np.zeros((3, 1))
    m["s"] = 0.0
    with pytest.raises(ValueError, match="singular"):
        G.predict_FRK(m, se_report=True)
    m["s"] = 1.0
    m["missing"] = {"miss": np.zeros((4, 1))}                     # pattern rows != picked rows
    with pytest.raises(ValueError, match="not compatible"):
        G.predict_FRK(m, se_report=True)


def test_out_of_scope_branches_are_loud():
    with pytest.raises(NotImplementedError, match="non-diagonal D"):                       # spam's sparse solve
        G.indeMLE(np.ones((5, 1)), np.ones((5, 3)), D=np.ones((5, 5)))
    with pytest.raises(NotImplementedError, match="LatticeKrig"):
        G.autoFRK(np.ones((5, 1)), np.zeros((5, 2)), finescale=True)
    with pytest.raises(NotImplementedError, match="LatticeKrig"):
        G.indeMLE(np.ones((5, 1)), np.ones((5, 3)), DfromLK={})


def test_default_K_grid():
    """K <- unique(round(seq(d + 1, maxK, by = maxK^(1/3) * d))), at most 30 values (R/autoFRK.R:254-257), worked by hand."""
    assert G.kgrid(2, 100).tolist() == [3, 12, 22, 31, 40, 49, 59, 68, 77, 87, 96]
    assert G.kgrid(1, 30).tolist() == [2, 5, 8, 11, 14, 18, 21, 24, 27, 30]
    big = G.kgrid(1, 4000)                                        # 252 candidates -> 30 equally spaced ones
    assert big.size == 30 and big[0] == 2 and big[-1] == 4000 and np.all(np.diff(big) > 0)
