"""GPU: parity at the shapes of the BASELINE.json configurations, at sizes the numpy oracle finishes in seconds
("parity tests never touch the BASELINE shapes").  Same tolerance as everywhere: 1e-8 relative per
component after sign alignment, census exact, trip counts exact where the oracle's exit has margin.

  C2  single-block PLS1, p >> n, KERNEL and SIMPLS, 10 components          (n=300 x p=6,000)
  C3  8 uneven blocks (ratios of 20k..450k), Y n x 10, NIPALS, 20 components (n=2,000 x p=3,000)
  C4  headline: 4 blocks 1:2:3:4, n=10,000, PLS1 NIPALS, 20 components, dense (p=4,000) and 10 % NaN (p=1,000)
  C5  tall n >> p, KERNEL and UNIPALS, 30 components, + predict              (n=20,000 x p=200)
The live-reference fixtures of the same configurations (tests/golden/live_c*_*.npz, smaller n) are checked by
test_gpu_nipals.py / test_gpu_methods.py.
"""
import warnings

import numpy as np
import pytest

from helpers import assert_trips, compare, snapshot_model

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _both(kw, X, Y, Xt, Yt, **rt):
    from mbpls_b200 import MBPLS
    from oracle import OracleMBPLS
    cp = (lambda a: [x.copy() for x in a] if isinstance(a, list) else a.copy())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = OracleMBPLS(**kw).fit(cp(X), cp(Y))
        m = MBPLS(**kw).set_runtime(**rt).fit(cp(X), cp(Y))
    ref, ours = snapshot_model(o, cp(Xt), cp(Yt)), snapshot_model(m, cp(Xt), cp(Yt))
    worst = compare(ours, ref, TOL, str(kw))
    if kw.get("method", "NIPALS") == "NIPALS":
        assert_trips(list(m.n_iter_), list(o.n_iter_), o.diff_trace_, kw.get("max_tol", 1e-14), str(kw))
    return m, o, worst


@pytest.mark.parametrize("method", ["KERNEL", "SIMPLS"])
def test_c2_pls1_wide_single_block(method):
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(300, (6000,), 1, 10, seed=31, noise=0.05, decay=0.8)
    Xt, Yt = latent_blocks(25, (6000,), 1, 10, seed=32, noise=0.05, decay=0.8)
    _both(dict(n_components=10, method=method, full_svd=True), X[0], Y.ravel(), Xt[0], Yt.ravel())


def test_c3_eight_uneven_blocks_pls2_q10():
    from oracle.cases import latent_blocks
    sizes = (50, 90, 150, 240, 350, 450, 550, 1120)  # the C3 block widths (20k ... 450k) scaled by 1/400
    X, Y = latent_blocks(2000, sizes, 10, 20, seed=33, noise=0.02, decay=0.85)
    Xt, Yt = latent_blocks(40, sizes, 10, 20, seed=34, noise=0.02, decay=0.85)
    m, o, _ = _both(dict(n_components=20, method="NIPALS"), X, Y, Xt, Yt)
    assert m.A_.shape == (8, 20) and np.allclose(m.A_.sum(axis=0), 1.0)
    assert sum(m.n_iter_) > 200  # a real PLS2 loop: tens of trips per component


@pytest.mark.parametrize("nan_frac,sizes", [(0.0, (400, 800, 1200, 1600)), (0.10, (100, 200, 300, 400))])
def test_c4_headline_shape_pls1_k20(nan_frac, sizes):
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(10_000, sizes, 1, 20, seed=35, noise=0.02, decay=0.85, nan_frac=nan_frac)
    Xt, Yt = latent_blocks(30, sizes, 1, 20, seed=36, noise=0.02, decay=0.85, nan_frac=nan_frac)
    m, o, _ = _both(dict(n_components=20, method="NIPALS", sparse_data=nan_frac > 0), X, Y, Xt, Yt)
    if nan_frac > 0:
        for b in range(4):
            for a, r in zip(m.sparse_X_info_[b], o.sparse_X_info_[b]):
                assert np.array_equal(a, r)
            assert np.array_equal(np.asarray(m.x_scalers_[b].n_samples_seen_), np.asarray(o.x_scalers_[b].n_samples_seen_))


@pytest.mark.parametrize("method", ["KERNEL", "UNIPALS"])
def test_c5_tall_n_much_larger_than_p_k30(method):
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(20_000, (100, 100), 4, 30, seed=37, noise=0.05, decay=0.9)
    Xt, Yt = latent_blocks(500, (100, 100), 4, 30, seed=38, noise=0.05, decay=0.9)
    m, o, _ = _both(dict(n_components=30, method=method, full_svd=True), X, Y, Xt, Yt)
    # batched predict on a fresh tall batch, through a device-resident (feature-major) input as the C5 bench does
    import torch
    Xb, _ = latent_blocks(20_000, (100, 100), 4, 30, seed=39, noise=0.05, decay=0.9)
    dev_blocks = [torch.from_numpy(np.asfortranarray(x)).cuda() for x in Xb]
    yh = m.predict(dev_blocks)
    ref = o.predict([x.copy() for x in Xb])
    assert np.linalg.norm(yh - ref) / np.linalg.norm(ref) < TOL
