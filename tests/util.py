"""Shared helpers for the parity tests."""
import numpy as np

TOL = 1e-10  # BASELINE.json north_star: FP64, 1e-10 relative (column-max-relative)


def colrel(a, b):
    """max over columns of  max_i |a - b| / max_i |b|  -- the error measure of SURVEY.md 8(c):
    elementwise-relative is meaningless near the zero crossings of a basis function."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    num = np.max(np.abs(a - b), axis=0)
    den = np.max(np.abs(b), axis=0)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(num / den))


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den else 1.0))


def make_fit(d, m, kstar, seed=0):
    """Random knots + oracle fit (UZ, BBBH, nconst), the inputs of mrtsrcpp_predict."""
    import oracle as O
    from oracle.mrts_oracle import xobs_diag_of

    rng = np.random.default_rng(seed)
    Xu = rng.random((m, d))
    xd = xobs_diag_of(Xu, m)
    fit = O.mrtsrcpp(Xu, xd, kstar)
    return Xu, xd, fit, rng


OBSERVED = {}   # label -> (largest observed error, bound): dumped to gpurun_out/observed_errors.json by conftest.py


def within(label, err, tol):
    """assert err < tol, remembering the observed error: bounds are meant to sit ~10x above what is observed on the
    hardware (profiles/observed_errors_r02.json lists them), not at a round number that merely passes."""
    err = float(err)
    old = OBSERVED.get(label)
    OBSERVED[label] = (max(err, old[0]) if old else err, tol)
    assert err < tol, (label, err, tol)
