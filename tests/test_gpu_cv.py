"""GPU: the device-resident cross-validation driver against an explicit refit loop with the oracle,
i.e. what sklearn.model_selection.cross_val_predict does with the reference estimator."""
import warnings

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _oracle_cv(kw, X, Y, folds):
    from oracle import OracleMBPLS
    out = np.full((Y.shape[0], Y.shape[1] if Y.ndim == 2 else 1), np.nan)
    for tr, te in folds:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = OracleMBPLS(**kw).fit([x[tr] for x in X], Y[tr])
            out[te] = m.predict([x[te] for x in X])
    return out


@pytest.mark.parametrize("method", ["NIPALS", "UNIPALS", "SIMPLS", "KERNEL"])
def test_leave_one_out_matches_refit_loop(method):
    from sklearn.model_selection import LeaveOneOut
    from mbpls_b200 import MBPLS
    from mbpls_b200.model_selection import cross_val_predict
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(24, (15, 9), 2, 3, seed=21)
    folds = list(LeaveOneOut().split(np.arange(24)))
    kw = dict(n_components=3, method=method, full_svd=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = cross_val_predict(MBPLS(**kw), X, Y, cv=LeaveOneOut())
    want = _oracle_cv(kw, X, Y, folds)
    assert got.shape == want.shape and rel_err(got, want) < 1e-9


@pytest.mark.parametrize("method", ["NIPALS", "UNIPALS", "SIMPLS", "KERNEL"])
def test_component_path_from_one_fit_per_fold(method):
    """Predictions for k = 1..K read off one K-component fit per fold equal K separate cross-validations."""
    from mbpls_b200 import MBPLS
    from mbpls_b200.model_selection import cross_val_predict
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(40, (30, 22), 1, 4, seed=22)
    y = Y.ravel()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        path = cross_val_predict(MBPLS(n_components=2, method=method, full_svd=True), X, y, cv=5, n_components_list=[1, 2, 3, 4])
    from sklearn.model_selection import KFold
    folds = list(KFold(n_splits=5).split(np.arange(40)))
    for k in (1, 2, 3, 4):
        want = _oracle_cv(dict(n_components=k, method=method, full_svd=True), X, Y, folds).ravel()
        assert path[k].shape == y.shape and rel_err(path[k], want) < 1e-9, (method, k)


def test_matches_sklearn_cross_val_predict_of_our_estimator():
    from sklearn.model_selection import cross_val_predict as sk_cvp
    from mbpls_b200 import MBPLS
    from mbpls_b200.model_selection import cross_val_predict
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(30, (12,), 1, 2, seed=23)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = cross_val_predict(MBPLS(n_components=2), X[0], Y.ravel(), cv=6)
        b = sk_cvp(MBPLS(n_components=2), X[0], Y.ravel(), cv=6)
    assert rel_err(a, np.asarray(b).ravel()) < 1e-10
