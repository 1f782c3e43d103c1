"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI by the
MBPLS host class, against the committed golden vectors and the live numpy oracle.

Tolerance: 1e-8 relative per component after sign alignment for scores / loadings / weights / beta
(north_star); NaN census, scaler observed counts and NIPALS trip counts exact.
"""
import warnings

import numpy as np
import pytest

from helpers import GOLDEN, assert_fixture_trips, assert_trips, compare, live_cases, load_live, rel_err, snapshot_model

pytestmark = pytest.mark.gpu

TOL = 1e-8
NIPALS_CASES = [c for c in live_cases() if "nipals" in c and "unipals" not in c]


def _fit(kwargs, X, Y, **rt):
    from mbpls_b200 import MBPLS
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MBPLS(**kwargs)
        if rt:
            m.set_runtime(**rt)
        m.fit([x.copy() for x in X] if isinstance(X, list) else X.copy(), Y.copy())
    return m


@pytest.mark.parametrize("name", NIPALS_CASES)
def test_nipals_matches_reference_fixture(name):
    X, Y, Xt, Yt, kwargs, ref = load_live(name)
    m = _fit(kwargs, X, Y)
    ours = snapshot_model(m, Xt, Yt)
    compare(ours, ref, TOL, name)
    assert_fixture_trips(ours["n_iter_"], name, ref)


def test_nipals_matches_reference_kat_csv():
    """mbpls/tests/test_mbpls.py:66-117 for method='NIPALS', with our estimator in place of the reference's."""
    import os
    z = np.load(os.path.join(GOLDEN, "kat_pn.npz"))
    X = [z["in/x1_train"], z["in/x2_train"]]
    Xt = [z["in/x1_test"], z["in/x2_test"]]
    m = _fit(dict(n_components=2, method="NIPALS", standardize=True, full_svd=True), X, z["in/y_train"])
    csv = lambda nm: z[f"csv/NIPALS/{nm}"]
    assert np.allclose(abs(np.concatenate(m.T_, axis=1)), abs(csv("T")))
    assert np.allclose(m.A_, csv("A"))
    Ts_t, T_t, U_t = m.transform(Xt, z["in/y_test"], return_block_scores=True)
    assert np.allclose(abs(np.concatenate(T_t, axis=1)), abs(csv("T_test")))
    assert np.allclose(abs(Ts_t), abs(csv("Ts_test")))
    assert np.allclose(abs(U_t), abs(csv("U_test")))
    assert np.allclose(abs(m.P_[0]), abs(csv("P1"))) and np.allclose(abs(m.P_[1]), abs(csv("P2")))
    assert np.allclose(abs(m.Ts_), abs(csv("Ts"))) and np.allclose(abs(m.U_), abs(csv("U")))
    assert np.allclose(abs(m.V_), abs(csv("V")))
    assert np.allclose(m.beta_, csv("beta"))
    assert np.allclose(m.predict(Xt), csv("Y_predict_test"))


@pytest.mark.parametrize("n,sizes,q,nan_frac", [
    (300, (700, 1300), 3, 0.0),     # short features: warp-per-feature resident pipeline
    (2500, (260, 400, 150), 4, 0.0),  # long features: CTA-wide resident pipeline
    (301, (333, 77), 1, 0.0),       # odd n, ragged sizes, PLS1
    (300, (700, 1300), 3, 0.10),    # NaN mode
    (2500, (260, 400), 2, 0.05),
])
def test_nipals_matches_live_oracle(n, sizes, q, nan_frac):
    from oracle import OracleMBPLS
    from oracle.cases import latent_blocks
    K = 4
    X, Y = latent_blocks(n, sizes, q, K, seed=n + len(sizes), nan_frac=nan_frac)
    Xt, Yt = latent_blocks(17, sizes, q, K, seed=5, nan_frac=nan_frac)
    kw = dict(n_components=K, method="NIPALS", sparse_data=nan_frac > 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(**kw).fit([x.copy() for x in X], Y.copy())
    ref = snapshot_model(o, Xt, Yt)
    m = _fit(kw, X, Y)
    ours = snapshot_model(m, Xt, Yt)
    compare(ours, ref, TOL, f"live n={n}")
    assert_trips(list(m.n_iter_), list(o.n_iter_), o.diff_trace_, 1e-14, f"live n={n}")


@pytest.mark.parametrize("rt", [dict(deflate_mode=1, standardize_mode=1), dict(fuse_next_xtu=False),
                                dict(trips_per_sync=1), dict(trips_per_sync=7), dict(materialize=False)])
def test_runtime_variants_agree(rt):
    """Fallback kernels, unfused first-XtU, and different readback batching all give the same model."""
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(500, (900, 300), 2, 3, seed=9)
    base = _fit(dict(n_components=3), X, Y)
    var = _fit(dict(n_components=3), X, Y, **rt)
    assert list(base.n_iter_) == list(var.n_iter_)
    assert rel_err(var.beta_, base.beta_) < 1e-12
    assert rel_err(var.Ts_, base.Ts_) < 1e-12
    for a, b in zip(var.P_, base.P_):
        assert rel_err(a, b) < 1e-12


@pytest.mark.parametrize("n", [1500, 2501, 9000])
def test_long_feature_kernel_variants_agree(n):
    """1024 < n <= 16384: register-resident with shared-memory ts/u0 (mode 0), global fallback (mode 1), the
    warp-specialised shared-memory bulk-copy pipeline (mode 2) and the L1-vector register variant (mode 3) must
    produce the same model; odd n exercises the padding."""
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(n, (150, 90), 2, 3, seed=n)
    fits = [_fit(dict(n_components=3), X, Y, deflate_mode=m, standardize_mode=min(m, 2)) for m in (0, 1, 2, 3)]
    for other in fits[1:]:
        assert list(other.n_iter_) == list(fits[0].n_iter_)
        assert rel_err(other.beta_, fits[0].beta_) < 1e-11
        assert rel_err(other.Ts_, fits[0].Ts_) < 1e-11
        assert rel_err(other.x_scalers_[0].scale_, fits[0].x_scalers_[0].scale_) < 1e-13


def test_standardize_kernel_matches_sklearn_semantics():
    import torch
    from mbpls_b200 import engine as E
    from oracle import OracleScaler
    rng = np.random.default_rng(1)
    for n in (57, 1500, 2049):
        X = rng.standard_normal((n, 40)) * rng.uniform(0.1, 30, 40) + rng.uniform(-4, 4, 40)
        X[:, 5] = 1.25
        X[:, 6] *= 1e-90   # scales far from 1: the quotients stay ordinary numbers
        X[:, 7] *= 1e+120
        X[rng.random(X.shape) < 0.07] = np.nan
        ref = OracleScaler().fit(X)
        for mode in (0, 1, 2):
            dev = torch.device("cuda:0")
            Xt = E.alloc_feature_major(40, n, dev)
            E.ingest_feature_major(X, n, 0, 40, Xt, dev)
            st = E.standardize_fit(Xt, n, mode)
            assert np.allclose(st.mean[:40].cpu().numpy(), ref.mean_, rtol=1e-13, atol=1e-15)
            assert np.allclose(st.var[:40].cpu().numpy(), ref.var_, rtol=1e-12, atol=1e-300)
            assert np.allclose(st.scale[:40].cpu().numpy(), ref.scale_, rtol=1e-12)
            assert np.array_equal(st.seen[:40].cpu().numpy(), np.broadcast_to(ref.n_samples_seen_, (40,)))
            Z = Xt[:, :n].cpu().numpy().T
            assert np.allclose(Z, ref.transform(X), rtol=1e-12, atol=1e-14, equal_nan=True)
            assert bool((Xt[:, n:] == 0).all())
            # the kernels divide by the scale through a reciprocal + FMA correction (csrc/common.cuh UniformDivisor), which is
            # claimed to be the correctly rounded quotient: with the device's own statistics numpy's true division must give
            # the same bits (also for the tiny / huge scales of columns 6 and 7)
            m_d, s_d = st.mean[:40].cpu().numpy(), st.scale[:40].cpu().numpy()
            want = (X - m_d) / s_d
            assert np.array_equal(np.nan_to_num(Z, nan=-7.0), np.nan_to_num(want, nan=-7.0))


def test_ingest_layouts_and_roundtrip():
    import torch
    from mbpls_b200 import engine as E
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(2)
    A = rng.standard_normal((45, 70))
    want = A.T
    for src in (A, np.asfortranarray(A), A[:, ::1], torch.from_numpy(A), torch.from_numpy(A).to(dev),
                torch.from_numpy(np.ascontiguousarray(A.T)).to(dev).t(), A.tolist()):
        Xt = E.alloc_feature_major(70, 45, dev)
        E.ingest_feature_major(src, 45, 0, 70, Xt, dev)
        assert np.array_equal(Xt[:, :45].cpu().numpy(), want)
    Xt = E.alloc_feature_major(30, 45, dev)
    E.ingest_feature_major(A, 45, 20, 50, Xt, dev)  # column sub-range (sharding)
    assert np.array_equal(Xt[:, :45].cpu().numpy(), A[:, 20:50].T)


def test_error_behaviour_matches_reference():
    from mbpls_b200 import MBPLS
    from sklearn.exceptions import NotFittedError
    X, y = np.random.rand(20, 6), np.random.rand(20)
    with pytest.raises(NameError):
        MBPLS(method="FOO").fit(X, y)
    with pytest.raises(ValueError):
        MBPLS().fit(X, y[:-1])
    Xn = X.copy()
    Xn[3, 2] = np.nan
    with pytest.raises(ValueError):
        MBPLS().fit(Xn, y)
    with pytest.raises(ValueError):
        MBPLS(standardize=False).fit(Xn, y)
    with pytest.raises(NotFittedError):
        MBPLS().predict(X)
    with pytest.warns(UserWarning):
        m = MBPLS(method="KERNEL", sparse_data=True, n_components=1)
        m.fit(Xn, y.reshape(-1, 1))
    assert m.method == "NIPALS"


def test_sklearn_surface():
    import pickle
    from sklearn.base import clone
    from mbpls_b200 import MBPLS
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(80, (30, 20), 1, 2, seed=4)
    m = MBPLS(n_components=2).fit(X, Y.ravel())
    assert set(m.get_params()) == {"n_components", "full_svd", "method", "standardize", "max_tol",
                                   "nipals_convergence_norm", "calc_all", "sparse_data", "copy"}
    c = clone(m).fit(X, Y.ravel())
    assert np.allclose(c.beta_, m.beta_, rtol=1e-12, atol=0)
    m2 = pickle.loads(pickle.dumps(m))
    assert np.allclose(m2.predict(X), m.predict(X), rtol=1e-12)
    assert 0.5 < m.score(X, Y.ravel()) <= 1.0
    assert np.isclose(m.r2_score(X, Y.ravel()), m.score(X, Y.ravel()))
    assert m.x_scalers_[0].inverse_transform(m.P_[0].T).shape == (2, 30)


def test_large_shape_invariants_and_device_inputs():
    """Size-independent properties at a shape the oracle cannot hold comfortably (n=10,000, p=60,000, device-generated,
    adopted zero-copy): orthonormal superscores, unit block weights, A_ columns sum to 1, beta_ = R_ V_', predict ==
    inverse_scale(Ts V_'), predict leaves a device input untouched, trips == 2 for PLS1."""
    import torch
    from mbpls_b200 import MBPLS, synth, engine as E
    dev = torch.device("cuda:0")
    n, sizes, K = 10_000, (10_000, 20_000, 30_000), 6
    p = sum(sizes)
    off = np.concatenate(([0], np.cumsum(sizes)))
    Xbuf = torch.empty((p, E.round_ld(n)), dtype=torch.float64, device=dev)

    def blocks():
        synth.fill_feature_major(Xbuf, n, 0, p, K, 77, noise=0.02, decay=0.85)
        return [Xbuf[off[b]:off[b + 1], :n].t() for b in range(len(sizes))]

    Y = synth.response(n, 1, K, dev, 77, decay=0.85)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MBPLS(n_components=K, copy=False).fit(blocks(), Y)
    assert m.n_iter_ == [2] * K
    G = m.Ts_.T @ m.Ts_
    assert np.allclose(G, np.eye(K), atol=1e-10)
    for Wb in m.W_:
        assert np.allclose(np.linalg.norm(Wb, axis=0), 1.0, rtol=1e-12)
    assert np.allclose(m.A_.sum(axis=0), 1.0, rtol=1e-12)
    assert np.allclose(m.beta_, m.R_ @ m.V_.T, rtol=1e-10, atol=1e-14)
    assert 0 < sum(m.explained_var_x_) <= 1 + 1e-9 and np.all(np.diff(np.cumsum(m.explained_var_y_)) >= -1e-12)
    Xnew = blocks()
    before = Xbuf[:, :n].clone()
    yh = m.predict(Xnew)
    assert torch.equal(Xbuf[:, :n], before), "predict must not modify a device input"
    Ts_new = m.transform(Xnew)
    yh2 = m.y_scaler_.inverse_transform(Ts_new @ m.V_.T)
    assert np.allclose(yh, yh2, rtol=1e-9, atol=1e-11)
    # the training data reproduce the training superscores
    assert np.allclose(np.abs(Ts_new / np.linalg.norm(Ts_new, axis=0)), np.abs(m.Ts_), rtol=1e-8, atol=1e-10)
    assert m.score(Xnew, Y.cpu().numpy()) > 0.9


def test_check_sparsity_level_matches_reference_semantics():
    from mbpls_b200 import MBPLS
    from oracle import nan_census
    rng = np.random.default_rng(0)
    A = rng.standard_normal((37, 53))
    A[rng.random(A.shape) < 0.03] = np.nan
    got = MBPLS().check_sparsity_level(A)
    want = nan_census(A)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
