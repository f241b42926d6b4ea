"""CPU oracle for the autoFRK hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import anything from this package.  The product package
(``autofrk_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference (egpivo/autoFRK v1.4.4) ships no tests, golden vectors or
fixtures for this path, R is not installed in the build container and
``src/autoFRK.cpp`` cannot be compiled here (Rcpp / RcppEigen / RSpectra headers are absent,
versions unpinned in DESCRIPTION:24).  The oracle is therefore a line-by-line restatement
(numpy for the dense algebra, plain C for the scalar kernel loops), each function citing the
reference file:line it follows, cross-checked by

  * the C restatement against the numpy restatement (two independent codings),
  * a 50-digit mpmath evaluation of the same formulas on small cases,
  * the algebraic identities the reference algorithm satisfies.

All ``file:line`` citations are relative to the upstream repository root.
"""
from .mrts_oracle import (  # noqa: F401
    tpm2, tpm_predict, mrts_core, mrtsrcpp, mrtsrcpp_predict0, mrtsrcpp_predict,
    mrts, predict_mrts, MrtsObject, fix_signs,
)
from .frk_oracle import (  # noqa: F401
    getASCeigens, getEigen, getHalf, sol_v, sol_eta, neg2llik, cMLE, cMLEimat, invCz, ZinvC,
    getLikelihood, selectBasis, indeMLE, autoFRK, predict_FRK, gram_stats, cmle_from_stats, cMLEsp, EM0miss, mkpd, ginv, knn_impute,
)
