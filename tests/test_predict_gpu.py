"""K1 parity on the GPU: fused TPS evaluation + projection vs the oracle, through the C ABI.

The fit (UZ, BBBH, nconst) comes from the oracle and is fed to both sides, which removes the
eigenvector sign/rotation ambiguity.  Tolerance: 1e-10 column-max-relative."""
import numpy as np
import pytest

from util import TOL, colrel, make_fit, within

pytestmark = pytest.mark.gpu

# (d, m, kstar, n2): cover every K1 slice configuration (BN 32..512, two slices), ragged rows / knots
CASES = [
    (2, 6, 2, 50),        # fewer knots than one chunk: the plan pads to 4 zero-weight chunks
    (1, 5, 3, 9),
    (3, 9, 4, 33),
    (2, 37, 4, 1),
    (2, 64, 20, 7),
    (2, 101, 28, 255),
    (2, 150, 60, 1000),
    (2, 200, 90, 1025),
    (2, 250, 125, 777),
    (2, 300, 160, 2049),
    (2, 500, 197, 4000),
    (2, 400, 250, 1500),
    (2, 520, 300, 3001),
    (2, 600, 380, 1999),
    (2, 1000, 497, 5000),
    (2, 801, 600, 1203),
    (3, 90, 30, 513),
    (3, 300, 100, 2000),
    (3, 700, 520, 1500),
    (1, 30, 8, 999),
    (1, 40, 10, 31),
    (2, 150, 3, 2500),      # BN 8 configuration (512-row tiles)
    (3, 120, 12, 1111),     # BN 16
    (2, 900, 380, 1500),    # BN 384, Phi three chunks ahead
    (2, 700, 310, 900),     # BN 320
]


@pytest.mark.parametrize("d,m,kstar,n2", CASES)
def test_mrtsrcpp_predict_matches_oracle(d, m, kstar, n2):
    import oracle as O
    from autofrk_b200 import mrts as G

    Xu, xd, fit, rng = make_fit(d, m, kstar, seed=m + kstar)
    xnew = rng.random((n2, d)) * 1.2 - 0.1
    nk = min(n2, 5)
    xnew[:nk] = Xu[:nk]                       # r == 0 hits (the `if (r != 0)` branch, src/autoFRK.cpp:137)
    K = kstar + d + 1                          # what predict.mrts passes (R/autoFRK.R:1458)
    ref = O.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], K)
    got = G.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], K)
    assert got["X1"].shape == (n2, K)
    assert np.all(got["X1"][:, kstar:] == 0.0)  # the d+1 trailing columns are exact zeros
    err = colrel(got["X1"][:, :kstar], ref["X1"][:, :kstar])
    within("test_mrtsrcpp_predict_matches_oracle:err", err, TOL)
    assert got["X"] is not None and np.array_equal(got["X"], Xu)  # quirk: X = Xu (src/autoFRK.cpp:320)
    # k = kstar (what mrtsrcpp_predict0 uses internally) gives the same columns
    got2 = G.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], kstar)
    assert np.array_equal(got2["X1"], got["X1"][:, :kstar])


def test_mrtsrcpp_predict_random_shapes_match_oracle():
    """24 seeded random (d, m, k*, n2): every K1 configuration and column-slice count the plan can pick, ragged row
    and column edges, against the oracle at 1e-10 (column-max-relative)."""
    import oracle as O
    from autofrk_b200 import mrts as G

    rs = np.random.default_rng(77)
    for it in range(24):
        d = int(rs.integers(1, 4))
        m = int(rs.integers(d + 6, 420))
        # d = 1 with many columns is intrinsically ill-conditioned (|gammas| ~ 1/rho_min): keep k* small there
        kmax = min(m - d - 1, 12 if d == 1 else 400)
        kstar = int(rs.integers(1, kmax + 1))
        n2 = int(rs.integers(1, 1500))
        Xu, xd, fit, rng = make_fit(d, m, kstar, seed=1000 + it)
        xnew = rng.random((n2, d)) * 1.2 - 0.1
        K = kstar + d + 1
        ref = O.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], K)["X1"]
        got = G.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], K)["X1"]
        assert got.shape == (n2, K) and np.all(got[:, kstar:] == 0.0)
        err = colrel(got[:, :kstar], ref[:, :kstar])
        within("test_mrtsrcpp_predict_random_shapes_match_oracle:err", err, TOL)


def test_predict_mrts_basis_and_identity():
    """predict.mrts end to end + the known-answer identity predict(mrts(knot,k), knot) == mrts(knot,k)."""
    import oracle as O
    from autofrk_b200 import mrts as G

    d, m, kstar = 2, 300, 120
    Xu, xd, fit, rng = make_fit(d, m, kstar, seed=5)
    K = kstar + d + 1
    oobj = O.MrtsObject(fit["X"], fit["UZ"], Xu, fit["nconst"], fit["BBBH"])
    gobj = G.MrtsBasis(fit["X"], fit["UZ"], Xu, fit["nconst"], fit["BBBH"])
    xnew = rng.random((1234, d))
    ref = O.predict_mrts(oobj, xnew)
    for use_plan in (True, False):
        got = G.predict_mrts(gobj, xnew, use_plan=use_plan)
        assert got.shape == (1234, K)
        within("test_predict_mrts_basis_and_identity:got", colrel(got, ref), TOL)
    atk = G.predict_mrts(gobj, Xu)
    within("test_predict_mrts_basis_and_identity:atk[", colrel(atk[:, d + 1:], fit["X"][:, d + 1:]), TOL)
    # column-subset object (selectBasis keeps the full UZ, R/autoFRK.R:324-328)
    sub = gobj.leading(40)
    refsub = O.predict_mrts(oobj.leading(40), xnew)
    within("test_predict_mrts_basis_and_identity:G.predict_mrtssub", colrel(G.predict_mrts(sub, xnew), refsub), TOL)


def test_empty_and_degenerate_inputs():
    from autofrk_b200 import mrts as G

    d, m, kstar = 2, 40, 10
    Xu, xd, fit, rng = make_fit(d, m, kstar, seed=3)
    out = G.mrtsrcpp_predict(Xu, xd, np.zeros((0, d)), fit["BBBH"], fit["UZ"], fit["nconst"], kstar)
    assert out["X1"].shape == (0, kstar)
    # all-zero eigenvector block: X1 is identically zero, no kernel launch needed
    UZ0 = np.zeros_like(fit["UZ"])
    out = G.mrtsrcpp_predict(Xu, xd, rng.random((17, d)), fit["BBBH"], UZ0, fit["nconst"], kstar)
    assert np.all(out["X1"] == 0.0)
    # non-finite coordinates propagate as NaN / stay confined to their own rows
    x = rng.random((64, d))
    x[5, 0] = np.nan
    got = G.mrtsrcpp_predict(Xu, xd, x, fit["BBBH"], fit["UZ"], fit["nconst"], kstar)["X1"]
    assert np.all(np.isnan(got[5])) and np.all(np.isfinite(np.delete(got, 5, axis=0)))
    # very large / very small coordinate scales (the table log and the MUFU-seeded sqrt take any normal double)
    import oracle as O
    for d2, scale in ((2, 1e-60), (2, 1e60), (3, 1e-100), (3, 1e100)):
        Xs = rng.random((30, d2)) * scale
        fit2 = O.mrtsrcpp(Xs, np.eye(d2), 6)
        xn = rng.random((40, d2)) * scale
        ref = O.mrtsrcpp_predict(Xs, None, xn, fit2["BBBH"], fit2["UZ"], fit2["nconst"], 6)["X1"]
        got = G.mrtsrcpp_predict(Xs, None, xn, fit2["BBBH"], fit2["UZ"], fit2["nconst"], 6)["X1"]
        assert colrel(got, ref) < 1e-9, (d2, scale, colrel(got, ref))


def test_error_messages_match_reference():
    from autofrk_b200 import mrts as G
    from autofrk_b200._lib import FrkError

    d, m, kstar = 2, 50, 10
    Xu, xd, fit, rng = make_fit(d, m, kstar, seed=1)
    xnew = rng.random((10, d))
    with pytest.raises(FrkError, match="xnew must have 2 columns"):
        G.mrtsrcpp_predict(Xu, xd, rng.random((10, 3)), fit["BBBH"], fit["UZ"], fit["nconst"], kstar)
    with pytest.raises(FrkError, match="BBBH must be 3 x 50"):
        G.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"][:, :-1], fit["UZ"], fit["nconst"], kstar)
    with pytest.raises(FrkError, match="UZ dimensions are inconsistent with n and k"):
        G.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"], 99)
    with pytest.raises(FrkError, match="nconst must have length 2"):
        G.mrtsrcpp_predict(Xu, xd, xnew, fit["BBBH"], fit["UZ"], fit["nconst"][:1], kstar)


def test_predict_dev_torch_and_linearity():
    """device-pointer path + a size-independent property at a larger n2: X1 is linear in W."""
    import torch
    from autofrk_b200 import mrts as G

    d, m, kstar = 2, 500, 197
    Xu, xd, fit, rng = make_fit(d, m, kstar, seed=2)
    n2 = 300_000
    g = torch.Generator(device="cuda").manual_seed(3)
    xt = torch.rand((d, n2), dtype=torch.float64, device="cuda", generator=g).T   # column-major (n2, d)
    plan = G.MrtsPlan(Xu, fit["BBBH"], fit["UZ"], fit["nconst"], kstar)
    out = plan.predict_dev(xt)
    torch.cuda.synchronize()
    host = plan.predict(xt.cpu().numpy())
    assert np.array_equal(out.cpu().numpy(), host)          # same kernel, device vs host-buffer path
    UZ2 = fit["UZ"].copy()
    UZ2[:m, :kstar] *= 2.0                                    # exact in FP64
    plan2 = G.MrtsPlan(Xu, fit["BBBH"], UZ2, fit["nconst"], kstar)
    out2 = plan2.predict_dev(xt)
    torch.cuda.synchronize()
    assert torch.equal(out2, out * 2.0)
    assert plan.launches > 0


def test_output_guard_bands_untouched():
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are hunted with guard bands: a padded
    leading dimension and spare columns pre-filled with a sentinel must survive every K1 configuration."""
    import torch
    from autofrk_b200 import mrts as G
    from autofrk_b200._lib import FRK_OUT_BASIS

    rng = np.random.default_rng(12)
    sentinel = -7.25
    for d, m, kstar, n2 in [(2, 90, 28, 1001), (2, 400, 197, 3333), (2, 700, 497, 2047), (3, 300, 600, 1500), (1, 50, 9, 77),
                            (2, 120, 5, 999), (3, 64, 8, 515), (2, 800, 380, 1027), (2, 700, 310, 640), (2, 20, 2, 3000)]:
        Xu = rng.random((m, d))
        UZ = np.zeros((m + d + 1, kstar + d + 1))
        UZ[:m, :kstar] = rng.standard_normal((m, kstar))
        K = kstar + d + 1
        plan = G.MrtsPlan(Xu, rng.standard_normal((d + 1, m)), UZ, np.ones(d), K)
        ld = n2 + 37
        xfull = torch.rand((d, ld), dtype=torch.float64, device="cuda")
        x = xfull.T[:n2]                                      # input with its own padded leading dimension
        buf = torch.full((K + 2, ld), sentinel, dtype=torch.float64, device="cuda")
        out = buf.T[:n2, 1:K + 1]                             # n2 x K window: guard rows below, guard columns on both sides
        for mode in (0, FRK_OUT_BASIS):
            buf.fill_(sentinel)
            plan.predict_dev(x, out, mode=mode)
            torch.cuda.synchronize()
            full = buf.T
            assert torch.all(full[n2:, :] == sentinel), (d, m, kstar, mode, "rows past n2 written")
            assert torch.all(full[:, 0] == sentinel) and torch.all(full[:, K + 1] == sentinel), "guard columns written"
            assert not torch.any(full[:n2, 1:K + 1] == sentinel)
