"""Generates tests/golden/mrts_small.{npz,json}.

The reference (egpivo/autoFRK) ships no golden vectors and cannot run in the build container (no R,
no Rcpp/Eigen/Spectra headers), so these fixtures are produced by the ORACLE restatement
(oracle/mrts_oracle.py) -- they pin the oracle against regressions and give the GPU tests a
committed, oracle-independent target; they do not pin the oracle to the reference (PARITY UNPINNED).

    python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402
from oracle.mrts_oracle import xobs_diag_of  # noqa: E402

CASES = [dict(tag="d1", d=1, m=30, kstar=8, n2=64), dict(tag="d2", d=2, m=120, kstar=40, n2=257),
         dict(tag="d3", d=3, m=90, kstar=30, n2=100)]


def main():
    out = {}
    rng = np.random.default_rng(20240101)
    for c in CASES:
        Xu = rng.random((c["m"], c["d"]))
        xnew = rng.random((c["n2"], c["d"]))
        xnew[:3] = Xu[:3]
        fit = O.mrtsrcpp(Xu, xobs_diag_of(Xu, c["m"]), c["kstar"])
        X1 = O.mrtsrcpp_predict(Xu, None, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], c["kstar"])["X1"]
        t = c["tag"]
        out.update({f"{t}_Xu": Xu, f"{t}_xnew": xnew, f"{t}_UZ": fit["UZ"], f"{t}_BBBH": fit["BBBH"],
                    f"{t}_nconst": fit["nconst"], f"{t}_X": fit["X"], f"{t}_X1": X1})
    here = Path(__file__).resolve().parent
    np.savez_compressed(here / "mrts_small.npz", **out)
    (here / "mrts_small.json").write_text(json.dumps({"generator": "tests/golden/make_golden.py (oracle)", "cases": CASES}, indent=1))


if __name__ == "__main__":
    main()
