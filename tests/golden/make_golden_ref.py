"""Generates tests/golden/ref_vectors.{npz,json} FROM THE REFERENCE'S OWN SOURCE.

The vectors are outputs of oracle/_ref/libautofrk_ref.so = /root/reference/src/autoFRK.cpp compiled unmodified
(oracle/Makefile target `ref`; Rcpp/Eigen/Spectra resolved to the stand-ins of oracle/stub/).  What they pin:
  * tpm2 / tpm_predict (src/autoFRK.cpp:60-153): the reference's scalar arithmetic, bit for bit;
  * mrts / mrtsrcpp / mrtsrcpp_predict0 / mrtsrcpp_predict / getASCeigens (:24-32, :160-326): the reference's
    sequence of operations, block offsets, checks and messages, with the stand-ins' dense algebra underneath
    (rounding-level differences from a real Eigen build; eigenvector signs arbitrary).
They cannot pin the R-level functions (no R interpreter exists in this image).

    python tests/golden/make_golden_ref.py        # needs /root/reference (build container only)
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import build_oracle as ob  # noqa: E402
from oracle import ref as R  # noqa: E402

KERNEL_CASES = [dict(tag="k1", d=1, m=40, n2=70), dict(tag="k2", d=2, m=40, n2=70), dict(tag="k3", d=3, m=40, n2=70)]
FIT_CASES = [dict(tag="d1", d=1, m=30, kstar=8, n2=64), dict(tag="d2", d=2, m=120, kstar=40, n2=257),
             dict(tag="d3", d=3, m=90, kstar=30, n2=100),
             dict(tag="rd", d=1, m=30, kstar=28, n2=200)]          # man/mrts.Rd:32-44: 30 equispaced knots, k = 30


def xobs_diag_of(Xu):
    """R/autoFRK.R:1080: diag(sqrt(n/(n-1))/apply(Xu,2,sd), d) -- an input of the .Call, formed by the caller."""
    n = Xu.shape[0]
    return np.diag(np.sqrt(n / (n - 1.0)) / np.std(Xu, axis=0, ddof=1)).reshape(Xu.shape[1], Xu.shape[1])


def errors():
    """Every Rcpp::stop site reachable through the entry points, with the message the reference produces."""
    rng = np.random.default_rng(5)
    Xu = rng.random((20, 2))
    xd = xobs_diag_of(Xu)
    xn = rng.random((7, 2))
    fit = R.mrtsrcpp(Xu, xd, 5)
    calls = {
        "tpm2_d": lambda: R.tpm2_shaped(Xu, np.zeros((20, 20)), 4),
        "tpm2_cols": lambda: R.tpm2_shaped(Xu[:, :1], np.zeros((20, 20)), 2),
        "tpm2_L": lambda: R.tpm2_shaped(Xu, np.zeros((19, 20)), 2),
        "tpm_predict_d": lambda: R.tpm_predict_shaped(xn, Xu, np.zeros((7, 20)), 0),
        "tpm_predict_cols": lambda: R.tpm_predict_shaped(xn[:, :1], Xu, np.zeros((7, 20)), 2),
        "tpm_predict_L": lambda: R.tpm_predict_shaped(xn, Xu, np.zeros((7, 19)), 2),
        "mrts_k_positive": lambda: R.mrtsrcpp(Xu, xd, 0),
        "mrts_n": lambda: R.mrtsrcpp(Xu[:1], xd, 1),
        "mrts_k_lt_n": lambda: R.mrtsrcpp(Xu, xd, 20),
        "mrtsrcpp_xobs": lambda: R.mrtsrcpp(Xu, xd[:1, :1], 5),
        "predict0_xobs": lambda: R.mrtsrcpp_predict0(Xu, np.eye(3), xn, 5),
        "predict0_xnew": lambda: R.mrtsrcpp_predict0(Xu, xd, rng.random((7, 3)), 5),
        "predict_xnew": lambda: R.mrtsrcpp_predict(Xu, xd, rng.random((7, 1)), fit["BBBH"], fit["UZ"], fit["nconst"], 8),
        "predict_BBBH": lambda: R.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"][:, :19], fit["UZ"], fit["nconst"], 8),
        "predict_UZ": lambda: R.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"], fit["UZ"], fit["nconst"], 9),
        "predict_nconst": lambda: R.mrtsrcpp_predict(Xu, xd, xn, fit["BBBH"], fit["UZ"], fit["nconst"][:1], 8),
    }
    out = {}
    for name, f in calls.items():
        try:
            f()
            out[name] = None
        except R.RefError as e:
            out[name] = str(e)
    return out


def main():
    assert ob.build_ref() is not None and ob.REF_SRC.exists(), "needs the upstream sources"
    out = {}
    rng = np.random.default_rng(20261020)
    for c in KERNEL_CASES:
        P = rng.random((c["m"], c["d"]))
        Pn = rng.random((c["n2"], c["d"]))
        Pn[:4] = P[:4]                                      # r == 0 pairs
        Pn[4] = P[4] + 1e-9                                 # a tiny distance
        t = c["tag"]
        out.update({f"{t}_P": P, f"{t}_Pnew": Pn, f"{t}_tpm_predict": R.tpm_predict(Pn, P, c["d"]),
                    f"{t}_tpm2_upper": R.tpm2(P, c["d"])})
    for c in FIT_CASES:
        d, m, ks = c["d"], c["m"], c["kstar"]
        Xu = np.linspace(0.0, 1.0, m).reshape(-1, 1) if c["tag"] == "rd" else rng.random((m, d))
        xnew = np.linspace(0.0, 1.0, c["n2"]).reshape(-1, 1) if c["tag"] == "rd" else rng.random((c["n2"], d))
        if c["tag"] != "rd":
            xnew[:3] = Xu[:3]
        xd = xobs_diag_of(Xu)
        fit = R.mrtsrcpp(Xu, xd, ks)
        K = ks + d + 1
        pk = R.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], K)       # k = K (R/autoFRK.R:1458)
        ps = R.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], ks)
        p0 = R.mrtsrcpp_predict0(Xu, xd, xnew, ks)
        assert np.array_equal(pk["X1"][:, :ks], ps["X1"]) and np.array_equal(pk["X"], Xu)
        t = c["tag"]
        out.update({f"{t}_Xu": Xu, f"{t}_xobs_diag": xd, f"{t}_xnew": xnew, f"{t}_X": fit["X"], f"{t}_UZ": fit["UZ"],
                    f"{t}_BBBH": fit["BBBH"], f"{t}_nconst": fit["nconst"], f"{t}_X1_kK": pk["X1"],
                    f"{t}_X1_predict0": p0["X1"], f"{t}_UZ_predict0": p0["UZ"]})
    A = rng.standard_normal((12, 20))
    A = A @ A.T
    e = R.getASCeigens(A)
    out.update({"eig_A": A, "eig_value": e["value"], "eig_vector": e["vector"]})
    here = Path(__file__).resolve().parent
    np.savez_compressed(here / "ref_vectors.npz", **out)
    meta = {"generator": "tests/golden/make_golden_ref.py", "source": "oracle/_ref/libautofrk_ref.so = upstream "
            "src/autoFRK.cpp (egpivo/autoFRK v1.4.4) compiled unmodified behind oracle/stub/", "kernel_cases": KERNEL_CASES,
            "fit_cases": FIT_CASES, "error_messages": errors()}
    (here / "ref_vectors.json").write_text(json.dumps(meta, indent=1))
    print("wrote", here / "ref_vectors.npz", (here / "ref_vectors.npz").stat().st_size, "bytes")
    print(json.dumps(meta["error_messages"], indent=1))


if __name__ == "__main__":
    main()
