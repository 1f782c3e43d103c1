"""GPU parity for method in {SIMPLS, UNIPALS, KERNEL} (mbpls/mbpls.py:995-1048, :384-574, :576-807): the CUDA path
against the reference's KAT CSVs, the committed live-reference fixtures and the live numpy oracle.  Tolerance
1e-8 relative per component after sign alignment."""
import os
import warnings

import numpy as np
import pytest

from helpers import GOLDEN, compare, live_cases, load_live, rel_err, snapshot_model

pytestmark = pytest.mark.gpu
TOL = 1e-8
CASES = [c for c in live_cases() if c.endswith(("_unipals", "_kernel", "_simpls")) or c.startswith(("kernel_", "unipals_"))]


def _fit(kwargs, X, Y, **rt):
    from mbpls_b200 import MBPLS
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MBPLS(**kwargs)
        if rt:
            m.set_runtime(**rt)
        m.fit([x.copy() for x in X] if isinstance(X, list) else X.copy(), Y.copy())
    return m


def _snapshot(m, Xt, Yt):
    """Like helpers.snapshot_model but tolerant of the cases where the reference cannot produce block scores."""
    from oracle.make_golden import snapshot
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        snap = snapshot(m)
        if m.method != "SIMPLS" and getattr(m, "calc_all", True):
            Ts, T, U = m.transform(Xt, Yt, return_block_scores=True)
            for b, arr in enumerate(T):
                snap[f"tr_T/{b}"] = np.asarray(arr)
        else:
            Ts, U = m.transform(Xt, Yt)
        snap["tr_Ts"], snap["tr_U"] = np.asarray(Ts), np.asarray(U)
        snap["predict"] = np.asarray(m.predict(Xt))
    return snap


@pytest.mark.parametrize("name", CASES)
def test_method_matches_reference_fixture(name):
    X, Y, Xt, Yt, kwargs, ref = load_live(name)
    m = _fit(kwargs, X, Y)
    ours = _snapshot(m, Xt, Yt)
    skip = [k for k in ref if k not in ours and k.startswith("tr_")]
    compare(ours, ref, TOL, name, skip=skip)


@pytest.mark.parametrize("route", ["gram", "stream"])
@pytest.mark.parametrize("name", ["pls2_ngtp_unipals", "c5_tall_unipals", "unipals_nostd"])
def test_unipals_tall_both_routes_match_reference_fixture(name, route):
    """UNIPALS with n >= p: the p-space route (X'X and X'Y deflated algebraically, scores recovered after the loop) and the
    reference's streaming form (per-component passes over the deflated X) against the same live-reference fixtures."""
    X, Y, Xt, Yt, kwargs, ref = load_live(name)
    m = _fit(kwargs, X, Y, unipals_route=route)
    ours = _snapshot(m, Xt, Yt)
    skip = [k for k in ref if k not in ours and k.startswith("tr_")]
    compare(ours, ref, TOL, f"{name} [{route}]", skip=skip)


@pytest.mark.parametrize("tag,methods", [("pn", ["UNIPALS", "KERNEL", "SIMPLS"]), ("np", ["UNIPALS", "KERNEL"])])
def test_methods_match_reference_kat_csvs(tag, methods):
    """mbpls/tests/test_mbpls.py:34-421 with our estimator in place of the reference's."""
    z = np.load(os.path.join(GOLDEN, f"kat_{tag}.npz"))
    X = [z["in/x1_train"], z["in/x2_train"]]
    Xt = [z["in/x1_test"], z["in/x2_test"]]
    preds = []
    for method in methods:
        m = _fit(dict(n_components=2, method=method, standardize=True, full_svd=True), X, z["in/y_train"])
        csv = lambda nm: z[f"csv/{method}/{nm}"]
        if method != "SIMPLS":
            assert np.allclose(abs(np.concatenate(m.T_, axis=1)), abs(csv("T")))
            assert np.allclose(m.A_, csv("A"))
            Ts_t, T_t, U_t = m.transform(Xt, z["in/y_test"], return_block_scores=True)
            assert np.allclose(abs(np.concatenate(T_t, axis=1)), abs(csv("T_test")))
        else:
            Ts_t, U_t = m.transform(Xt, z["in/y_test"])
        assert np.allclose(abs(Ts_t), abs(csv("Ts_test")))
        assert np.allclose(abs(U_t), abs(csv("U_test")))
        assert np.allclose(abs(m.P_[0]), abs(csv("P1"))) and np.allclose(abs(m.P_[1]), abs(csv("P2")))
        assert np.allclose(abs(m.Ts_), abs(csv("Ts"))) and np.allclose(abs(m.U_), abs(csv("U")))
        assert np.allclose(abs(m.V_), abs(csv("V")))
        assert np.allclose(m.beta_, csv("beta"))
        pred = m.predict(Xt)
        assert np.allclose(pred, csv("Y_predict_test"))
        preds.append(pred)
    for pr in preds:
        assert np.allclose(preds[0], pr, atol=1e-3)  # test_mbpls.py:122-123


@pytest.mark.parametrize("method", ["SIMPLS", "UNIPALS", "KERNEL"])
@pytest.mark.parametrize("n,sizes,q", [(700, (300, 500), 3), (300, (900, 600, 77), 2), (1201, (257,), 1)])
def test_method_matches_live_oracle(method, n, sizes, q):
    from oracle import OracleMBPLS
    from oracle.cases import latent_blocks
    K = 4
    X, Y = latent_blocks(n, sizes, q, K, seed=n + q)
    Xt, Yt = latent_blocks(19, sizes, q, K, seed=7)
    kw = dict(n_components=K, method=method, full_svd=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(**kw).fit([x.copy() for x in X], Y.copy())
    ref = snapshot_model(o, Xt, Yt)
    m = _fit(kw, X, Y)
    ours = snapshot_model(m, Xt, Yt)
    compare(ours, ref, TOL, f"{method} n={n}")


def test_crossprod_dmma_matches_numpy():
    import torch
    from mbpls_b200 import crossmethods as CM, engine as E
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    for (p, n) in ((130, 1000), (257, 77), (64, 3001)):
        A = rng.standard_normal((p, n))
        Xt = E.alloc_feature_major(p, n, dev)
        Xt[:, :n] = torch.from_numpy(A).to(dev)
        ldv = (p + 15) // 16 * 16
        G = CM.crossprod(Xt, Xt, p, p, n, True, ldv)[:, :p].cpu().numpy()
        assert rel_err(G, A @ A.T) < 1e-13
        H = CM.crossprod(Xt, Xt, n, n, p, False, Xt.shape[1])[:, :n].cpu().numpy()
        assert rel_err(H, A.T @ A) < 1e-13


@pytest.mark.parametrize("m", [1, 2, 3, 7, 20, 33, 64])
def test_small_device_linear_algebra_matches_numpy(m):
    """One-sided Jacobi kernels (csrc/smalllin.cu) against numpy: top eigenvector of a PSD matrix, pinv of a general
    (also rank-deficient / ill-conditioned) matrix, top left singular vector of A B' from the two Gram matrices."""
    import torch
    from mbpls_b200 import engine as E
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(m)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    C = rng.standard_normal((m + 5, m)) * (0.5 ** np.arange(m))
    G = C.T @ C
    v = E.small_top_eigvec(t(G)).cpu().numpy()
    ref = np.linalg.svd(G)[0][:, 0]
    assert min(rel_err(v, ref), rel_err(-v, ref)) < 1e-10
    M = rng.standard_normal((m, m)) + np.triu(np.ones((m, m)))
    assert rel_err(E.small_pinv(t(M)).cpu().numpy(), np.linalg.pinv(M)) < 1e-9
    if m > 2:  # rank deficient
        Md = M.copy()
        Md[:, -1] = Md[:, 0] + Md[:, 1]
        assert rel_err(E.small_pinv(t(Md)).cpu().numpy(), np.linalg.pinv(Md)) < 1e-8
    A, Bm = rng.standard_normal((40 + m, m)), rng.standard_normal((40 + m, m))
    c = E.small_top_sv_product(t(Bm.T @ Bm), t(A.T @ A)).cpu().numpy()
    u = A @ c
    u /= np.linalg.norm(u)
    uref = np.linalg.svd(A @ Bm.T)[0][:, 0]
    assert min(rel_err(u, uref), rel_err(-u, uref)) < 1e-9


@pytest.mark.parametrize("method", ["KERNEL", "UNIPALS", "SIMPLS"])
def test_few_long_features_cut_the_sample_axis(method):
    """n >= 2^16 with a handful of features: X'Y / X'U / X'Ts cut every feature into sample chunks (mbpls_xt_multi_split_f64,
    partial products added in chunk order) and the n-space products of the fit go through the tall ring kernel; against the
    oracle on an odd n."""
    from oracle.cases import latent_blocks
    from oracle import OracleMBPLS
    from mbpls_b200 import MBPLS
    n = (1 << 18) + 4099
    X, Y = latent_blocks(n, (14, 10), 2, 3, seed=23)
    kw = dict(n_components=3, method=method)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(**kw, **({"full_svd": True} if method == "SIMPLS" else {})).fit([x.copy() for x in X], Y.copy())
        m = MBPLS(**kw).fit([x.copy() for x in X], Y.copy())
    ref, ours = snapshot_model(o, [x.copy() for x in X], Y.copy()), snapshot_model(m, [x.copy() for x in X], Y.copy())
    compare(ours, ref, 1e-8, f"{method} n={n}")
