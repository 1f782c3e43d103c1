"""TEST INFRASTRUCTURE ONLY -- seeded inputs shared by the golden generator, the tests and bench.py.

``kat_inputs`` regenerates the inputs of the reference's four known-answer tests
(``mbpls/tests/test_mbpls.py:36-57`` for P>N, ``:220-249`` for N>P) with the
historical ``ortho_group`` sampler, so that the reference's CSVs apply.
``latent_blocks`` is the latent-structure generator of SURVEY.md section 8(d).
"""
from __future__ import annotations

import numpy as np

from .mbpls_oracle import legacy_ortho_group_rvs


def kat_inputs(num_samples: int):
    """Inputs of the reference KATs: num_samples=50 (P>N) or 150 (N>P)."""
    rand_seed, p1, p2, noise = 25, 25, 45, 5
    np.random.seed(rand_seed)
    loading1 = np.expand_dims(np.random.randint(0, 10, p1), 1)
    loading2 = np.expand_dims(np.sin(np.linspace(0, 5, p2)), 1)
    y = legacy_ortho_group_rvs(num_samples, rand_seed)[:, :2]
    x1 = np.dot(y[:, 0:1], loading1.T)
    x2 = np.dot(y[:, 1:2], loading2.T)
    x1 = np.random.normal(x1, 0.05 * noise)
    x2 = np.random.normal(x2, 0.05 * noise)
    idx = np.random.choice(np.arange(num_samples), num_samples, replace=False)
    cut = round(num_samples * 8 / 10)
    tr, te = idx[:cut], idx[cut:]
    return dict(x1_train=x1[tr], x2_train=x2[tr], y_train=y[tr],
                x1_test=x1[te], x2_test=x2[te], y_test=y[te])


def latent_blocks(n, sizes, q, n_components, seed, noise=0.1, decay=0.7, nan_frac=0.0):
    """Low-rank-plus-noise blocks: X_b = Z L_b diag(s) + noise*E_b, Y = Z[:, :q] C + 0.05 E_Y.

    r = n_components + 5 latent factors with a geometric singular profile
    ``s_j = decay**j`` so NIPALS converges in tens of trips.
    ``nan_frac`` > 0 punches i.i.d. Bernoulli NaN holes into the X blocks only.
    """
    rng = np.random.default_rng(seed)
    r = n_components + 5
    Z = rng.standard_normal((n, r))
    s = decay ** np.arange(r)
    blocks = []
    for pb in sizes:
        L = rng.standard_normal((r, pb))
        Xb = (Z * s) @ L + noise * rng.standard_normal((n, pb))
        blocks.append(Xb)
    C = rng.standard_normal((min(q, r), q))
    Y = Z[:, :min(q, r)] @ C + 0.05 * rng.standard_normal((n, q))
    if nan_frac > 0:
        mrng = np.random.default_rng(seed + 1)
        for Xb in blocks:
            Xb[mrng.random(Xb.shape) < nan_frac] = np.nan
    return blocks, Y


def readme_quickstart(seed=0):
    """BASELINE config C1: ``README.rst:82-95`` (two uniform-random blocks, y 40x1, K=3)."""
    rng = np.random.RandomState(seed)
    x1 = rng.rand(40, 200)
    x2 = rng.rand(40, 250)
    y = rng.rand(40, 1)
    return [x1, x2], y
