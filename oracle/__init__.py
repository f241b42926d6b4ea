"""CPU oracle for the autoFRK hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import anything from this package.  The product package
(``autofrk_b200``) never imports it and has no CPU fallback.

PARITY: pinned to the reference's own source for everything ``src/autoFRK.cpp`` holds, unpinned for the R code.
The reference (egpivo/autoFRK v1.4.4) ships no tests, golden vectors or fixtures for this path
and R is not installed in the build container.  Its C++ translation unit IS compiled here, unmodified, behind stand-in
Rcpp / Eigen / Spectra headers (``oracle/stub/``, ``oracle/ref_wrap.cpp`` -> ``oracle/_ref/libautofrk_ref.so``, bound by
``oracle/ref.py``); ``tests/test_oracle_ref.py`` checks this restatement against it and against the vectors it
generated (``tests/golden/ref_vectors.npz``): the scalar kernel loops bit for bit, the fit / predict wrappers to
rounding level, every argument-check message verbatim.  The R-level functions (``frk_oracle.py``, the front-ends in
``mrts_oracle.py``) have no executable reference here; they are a line-by-line restatement (numpy for the dense algebra),
each function citing the reference file:line it follows, cross-checked by

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
