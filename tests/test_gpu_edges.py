"""GPU: edge cases of the fit path against the live oracle -- ragged / tiny shapes, single-feature blocks, one
component, constant columns (scale 1 rule), K equal to the rank budget, many blocks, float32 / list inputs,
NaN mode with a fully observed block, copy=False semantics."""
import os
import warnings

import numpy as np
import pytest

from helpers import assert_trips, compare, rel_err, snapshot_model

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _both(kw, X, Y, Xt=None, Yt=None):
    from mbpls_b200 import MBPLS
    from oracle import OracleMBPLS
    cp = (lambda a: [np.array(x, dtype=float) for x in a] if isinstance(a, list) else np.array(a, dtype=float))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(**kw).fit(cp(X), np.array(Y, dtype=float))
        m = MBPLS(**kw).fit(cp(X), np.array(Y, dtype=float))
    if Xt is None:
        Xt, Yt = X, Y
    ref, ours = snapshot_model(o, cp(Xt), np.array(Yt, dtype=float)), snapshot_model(m, cp(Xt), np.array(Yt, dtype=float))
    compare(ours, ref, TOL, str(kw))
    if kw.get("method", "NIPALS") == "NIPALS":
        assert_trips(list(m.n_iter_), list(o.n_iter_), o.diff_trace_, kw.get("max_tol", 1e-14))
    return m, o


@pytest.mark.parametrize("method", ["NIPALS", "UNIPALS", "KERNEL", "SIMPLS"])
def test_tiny_and_ragged_shapes(method):
    rng = np.random.default_rng(1)
    Z = rng.standard_normal((9, 3))
    X = [Z @ rng.standard_normal((3, 1)) + 0.1 * rng.standard_normal((9, 1)),      # single-feature block
         Z @ rng.standard_normal((3, 5)) + 0.1 * rng.standard_normal((9, 5)),
         Z @ rng.standard_normal((3, 2)) + 0.1 * rng.standard_normal((9, 2))]
    Y = Z[:, :2] @ rng.standard_normal((2, 2)) + 0.05 * rng.standard_normal((9, 2))
    _both(dict(n_components=2, method=method, full_svd=True), X, Y)


@pytest.mark.parametrize("method", ["NIPALS", "UNIPALS", "KERNEL", "SIMPLS"])
def test_single_block_single_component(method):
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(33, (17,), 1, 2, seed=3)
    m, o = _both(dict(n_components=1, method=method, full_svd=True), X[0], Y.ravel())
    if method != "SIMPLS":
        assert np.allclose(m.A_corrected_, 1.0) and m.A_corrected_.shape == (1, 1)  # mbpls.py:956-957


def test_constant_columns_get_unit_scale():
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(40, (12, 9), 2, 2, seed=4)
    X[0][:, 3] = 7.5       # constant feature -> scale_ == 1, standardised column == 0
    X[1][:, 0] = -2.0
    m, o = _both(dict(n_components=2), X, Y)
    assert m.x_scalers_[0].scale_[3] == 1.0 and m.x_scalers_[1].scale_[0] == 1.0
    assert np.all(m.W_[0][3] == 0.0)


def test_many_blocks_and_list_float32_inputs():
    rng = np.random.default_rng(5)
    Z = rng.standard_normal((50, 6))
    sizes = [3, 11, 1, 8, 20, 2, 5, 9, 4]
    X = [(Z @ rng.standard_normal((6, s)) + 0.05 * rng.standard_normal((50, s))).astype(np.float32) for s in sizes]
    Y = (Z[:, :3] @ rng.standard_normal((3, 3)) + 0.05 * rng.standard_normal((50, 3)))
    _both(dict(n_components=4), [x.astype(np.float64) for x in X], Y)
    from mbpls_b200 import MBPLS
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = MBPLS(n_components=3).fit([x.tolist() for x in X][0:1][0], Y[:, 0].tolist())  # nested lists -> single block
        b = MBPLS(n_components=3).fit(np.asarray(X[0], dtype=np.float64), Y[:, 0])
    assert rel_err(a.beta_, b.beta_) < 1e-12


def test_nan_mode_with_a_fully_observed_block_and_dense_rows():
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(70, (20, 35, 15), 2, 3, seed=6)
    rng = np.random.default_rng(7)
    X[1][rng.random(X[1].shape) < 0.05] = np.nan        # block 0 and 2 stay fully observed
    Xt, Yt = latent_blocks(9, (20, 35, 15), 2, 3, seed=8)
    Xt[2][0, 4] = np.nan
    m, o = _both(dict(n_components=3, sparse_data=True), X, Y, Xt, Yt)
    assert len(m.sparse_X_info_[0][0]) == 0 and len(m.sparse_X_info_[1][0]) > 0


def test_max_tol_and_norm_variants_change_trip_counts_like_the_reference():
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(80, (30, 50), 3, 3, seed=9)
    for kw in (dict(max_tol=1e-6), dict(max_tol=1e-10, nipals_convergence_norm=1), dict(max_tol=1e-9, nipals_convergence_norm=-np.inf),
               dict(max_tol=1e-8, nipals_convergence_norm="fro")):
        m, o = _both(dict(n_components=3, **kw), X, Y)
        assert list(m.n_iter_) == list(o.n_iter_)
    from mbpls_b200 import MBPLS
    with pytest.raises(ValueError):
        MBPLS(nipals_convergence_norm=0).fit(X, Y)


def test_copy_false_leaves_host_arrays_usable_and_model_identical():
    from mbpls_b200 import MBPLS
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(60, (25, 40), 2, 2, seed=10)
    keep = [x.copy() for x in X]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = MBPLS(n_components=2, copy=True).fit(X, Y)
        assert all(np.array_equal(x, k) for x, k in zip(X, keep))  # copy=True never touches the inputs (mbpls.py:303)
        b = MBPLS(n_components=2, copy=False).fit([x.copy() for x in X], Y.copy())
    assert rel_err(a.beta_, b.beta_) < 1e-13


def test_components_up_to_rank_budget():
    """n_components close to min(n, p): late components are tiny but every attribute must still match."""
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(14, (6, 5), 1, 2, seed=11, noise=0.3)
    _both(dict(n_components=6, method="SIMPLS", full_svd=True), X, Y.ravel())
    _both(dict(n_components=6, method="KERNEL", full_svd=True), X, Y.ravel())


def test_infinity_is_rejected_in_every_mode():
    """check_array semantics (mbpls.py:310,336): +-inf is an error even when NaN is allowed (sparse_data=True)."""
    from mbpls_b200 import MBPLS
    rng = np.random.default_rng(12)
    X, y = rng.standard_normal((30, 8)), rng.standard_normal((30, 1))
    Xi = X.copy()
    Xi[4, 2] = np.inf
    for kw in (dict(), dict(standardize=False), dict(sparse_data=True), dict(sparse_data=True, standardize=False)):
        with pytest.raises(ValueError), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            MBPLS(n_components=1, **kw).fit(Xi, y)
    m = MBPLS(n_components=1).fit(X, y)
    with pytest.raises(ValueError):
        m.predict(Xi)


def test_pandas_dataframes_and_dense_block_score_fast_path():
    """DataFrame blocks (what the reference's notebooks pass) and the one-pass block-score transform."""
    import pandas as pd
    from mbpls_b200 import MBPLS
    from oracle import OracleMBPLS
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(90, (23, 41, 12), 2, 5, seed=13)
    Xn, Yn = latent_blocks(31, (23, 41, 12), 2, 5, seed=14)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(n_components=5).fit([x.copy() for x in X], Y.copy())
        m = MBPLS(n_components=5).fit([pd.DataFrame(x) for x in X], pd.DataFrame(Y))
        for std in (True, False):
            if not std:
                o = OracleMBPLS(n_components=5, standardize=False).fit([x.copy() for x in X], Y.copy())
                m = MBPLS(n_components=5, standardize=False).fit([x.copy() for x in X], Y.copy())
            Ts_o, T_o = o.transform([x.copy() for x in Xn], return_block_scores=True)
            Ts_m, T_m = m.transform([pd.DataFrame(x) for x in Xn], return_block_scores=True)
            assert rel_err(np.abs(Ts_m), np.abs(Ts_o)) < 1e-9
            for a, b in zip(T_m, T_o):
                assert a.shape == b.shape and rel_err(np.abs(a), np.abs(b)) < 1e-9


def test_plot_data_matches_reference_recipe(capsys):
    """plot (mbpls.py:1439-1556) draws inverse-scaled loadings, block scores and importances; plot_data hands out
    exactly those arrays, computed here the way the reference's plot body does from the oracle's attributes."""
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(50, (14, 9), 2, 3, seed=8)
    m, o = _both(dict(n_components=3), X, Y)
    data = m.plot_data([1, 3])                                   # 1-based component numbers (:1459-1461)
    assert [d["component"] for d in data] == [1, 3]
    for d in data:
        k = d["component"] - 1
        s = np.sign(np.dot(m.Ts_[:, k], o.Ts_[:, k]))
        assert rel_err(d["importance_percent"], 100 * o.A_[:, k]) < TOL
        assert abs(d["explained_var_y_percent"] - 100 * o.explained_var_y_[k]) < 1e-6
        for b in range(2):
            # inverse_transform is affine, so the sign ambiguity of the component shows up around the mean
            want = o.x_scalers_[b].inverse_transform((s * o.P_[b][:, k]).reshape(1, -1)).ravel()
            assert rel_err(d["loadings"][b], want) < TOL
            assert rel_err(d["block_scores"][b], s * o.T_[b][:, k]) < TOL
    assert len(m.plot_data(2)) == 2                              # int: the first N components (:1455)
    assert len(m.plot_data(7)) == 3                              # truncated with the reference's printed note
    assert "shortened" in capsys.readouterr().out
    with pytest.raises(ValueError):
        m.plot_data([4])
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            m.plot(1)


@pytest.mark.parametrize("n_train,N", [(112, 160), (101, 160), (96, 96)])
def test_row_slice_views_of_a_taller_device_matrix_are_never_written(n_train, N):
    """A row slice ``X_cm[:n_train]`` of a column-major CUDA matrix has the parent's row count as its column stride: the
    elements [n_train, N) of every feature are the caller's held-out samples, not padding.  Neither predict / transform
    (which adopt such views zero-copy) nor fit(copy=False) may touch them, and a slice that starts at a later row must take
    the copy path instead of failing (engine.try_adopt_feature_major)."""
    import torch
    from mbpls_b200 import MBPLS
    from oracle import OracleMBPLS
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(N, (24, 40), 2, 3, seed=9)
    dev = torch.device("cuda:0")
    buf = torch.zeros((64, N), dtype=torch.float64, device=dev)      # feature-major parent: 64 features x N samples
    buf[:24] = torch.from_numpy(X[0].T.copy()).to(dev)
    buf[24:] = torch.from_numpy(X[1].T.copy()).to(dev)
    parent = buf.clone()
    views = lambda r0, r1: [buf[:24, r0:r1].t(), buf[24:, r0:r1].t()]  # n x p_b views with strides (1, N)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(n_components=3).fit([x[:n_train].copy() for x in X], Y[:n_train].copy())
        m = MBPLS(n_components=3).fit([x[:n_train].copy() for x in X], Y[:n_train].copy())
        yh = m.predict(views(0, n_train))
        assert torch.equal(buf, parent), "predict wrote into the caller's matrix"
        assert rel_err(yh, o.predict([x[:n_train] for x in X])) < TOL
        if N > n_train:
            yh_tail = m.predict(views(n_train, N))                    # slice with a storage offset
            assert torch.equal(buf, parent)
            assert rel_err(yh_tail, o.predict([x[n_train:] for x in X])) < TOL
        Ts = m.transform(views(0, n_train))
        assert torch.equal(buf, parent)
        assert rel_err(np.abs(Ts), np.abs(o.transform([x[:n_train] for x in X]))) < TOL
        m2 = MBPLS(n_components=3, copy=False).fit(views(0, n_train), Y[:n_train].copy())
        assert torch.equal(buf[:, n_train:], parent[:, n_train:]), "fit(copy=False) wrote beyond the n samples it was given"
        assert rel_err(m2.beta_, o.beta_) < TOL


def test_tall_batches_take_the_tma_ring_product_kernel():
    """predict / transform on m >= 2^18 new samples with at most 4 outputs run through skinny_tall_kernel (persistent CTAs fed
    by a ring of bulk copies, csrc/finalize.cu); same numbers as the reference's X.dot(beta_) / X.dot(R_) (:1386, :1117),
    ragged tail tile, odd m, NaN rows counted as zero (:1379-1383), +-inf rejected."""
    from mbpls_b200 import MBPLS, engine
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(150, (25, 14), 2, 3, seed=71)
    m = engine.TALL_MIN_SAMPLES + 4097  # not a multiple of the 2,048-sample tile, odd
    rng = np.random.default_rng(72)
    Xn = [rng.standard_normal((m, 25)), rng.standard_normal((m, 14))]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = MBPLS(n_components=3).fit([x.copy() for x in X], Y.copy())
        yh = model.predict(Xn)
        Ts = model.transform(Xn)
    Z = np.hstack([sc.transform(x) for sc, x in zip(model.x_scalers_, Xn)])
    want = model.y_scaler_.inverse_transform(Z @ model.beta_)
    assert yh.shape == want.shape and rel_err(yh, want) < 1e-12
    assert rel_err(Ts, Z @ model.R_) < 1e-12
    os.environ["MBPLS_TALL"] = "0"  # the per-thread-load kernel on the same batch
    try:
        assert rel_err(model.predict(Xn), want) < 1e-12
    finally:
        del os.environ["MBPLS_TALL"]
    # more than four outputs: the ring kernel makes one pass per four (superscores of a 6-component model)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model6 = MBPLS(n_components=6).fit([x.copy() for x in X], Y.copy())
        Ts6 = model6.transform(Xn)
    Z6 = np.hstack([sc.transform(x) for sc, x in zip(model6.x_scalers_, Xn)])
    assert Ts6.shape == (m, 6) and rel_err(Ts6, Z6 @ model6.R_) < 1e-12
    # NaN mode: missing entries of new data count as zero after scaling
    Xm = [x.copy() for x in Xn]
    Xm[0][::7, 3] = np.nan
    Xm[1][m - 1, 0] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sp = MBPLS(n_components=3, sparse_data=True).fit([x.copy() for x in X], Y.copy())
        got = sp.predict(Xm)
    Zm = np.nan_to_num(np.hstack([sc.transform(x) for sc, x in zip(sp.x_scalers_, Xm)]), nan=0.0)
    assert rel_err(got, sp.y_scaler_.inverse_transform(Zm @ sp.beta_)) < 1e-12
    with pytest.raises(ValueError):
        model.predict(Xm)  # dense model: NaN in new data is an error (check_array, :1368)
    Xi = [x.copy() for x in Xn]
    Xi[1][m - 2, 7] = np.inf
    with pytest.raises(ValueError):
        model.predict(Xi)


@pytest.mark.parametrize("method", ["NIPALS", "KERNEL", "SIMPLS", "UNIPALS"])
def test_unmaterialised_model_is_freed_without_the_cyclic_collector(method):
    """materialize=False keeps the fitted attributes as closures over device buffers; they must not form a reference cycle
    with the model (model -> _lazy -> closure -> model): a cycle frees hundreds of MB of device memory only when Python's
    cyclic collector happens to run, and the irregular frees sent PyTorch's caching allocator back to cudaMalloc in the middle
    of multi-GPU benchmark loops (60-90 ms per call with peer mappings)."""
    import gc
    import weakref
    from oracle.cases import latent_blocks
    from mbpls_b200 import MBPLS
    X, Y = latent_blocks(300, (700, 500), 2, 3, seed=5)
    gc.collect()
    gc.disable()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = MBPLS(n_components=3, method=method).set_runtime(materialize=False, small_path=False).fit([x.copy() for x in X], Y.copy())
        ref = weakref.ref(m)
        del m
        assert ref() is None
    finally:
        gc.enable()
