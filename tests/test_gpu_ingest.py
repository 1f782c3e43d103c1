"""GPU: the ingest path -- float32 sources widened on the device, pandas DataFrames, pickled blocks staged
in page-locked memory (mbpls_b200/ingest.py; the storage format of the reference's mbpls/data/**/*.pkl), read-only arrays."""
import os
import warnings

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _fit(X, Y, **kw):
    from mbpls_b200 import MBPLS
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return MBPLS(n_components=3, **kw).fit(X, Y)


@pytest.mark.parametrize("n,sizes", [(50, (30, 22)), (9000, (70, 45))])
def test_float32_sources_are_widened_on_the_device(n, sizes):
    """check_array(dtype=float64) (mbpls.py:310) widens float32 on the host; here the block crosses PCIe as float32 and is
    widened inside the transposition kernel: same bits as the host conversion."""
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(n, sizes, 2, 3, seed=61)
    X32 = [x.astype(np.float32) for x in X]
    a = _fit([x.copy() for x in X32], Y.copy())
    b = _fit([x.astype(np.float64) for x in X32], Y.copy())
    assert np.array_equal(a.beta_, b.beta_) and np.array_equal(a.Ts_, b.Ts_)
    import torch
    c = _fit([torch.from_numpy(x.copy()).cuda() for x in X32], Y.copy())  # float32 CUDA tensors
    assert rel_err(c.beta_, b.beta_) < 1e-12
    assert rel_err(a.predict([x.copy() for x in X32]), b.predict([x.astype(np.float64) for x in X32])) < 1e-12


def test_pandas_frames_pickled_blocks_and_pinned_staging(tmp_path):
    import pandas as pd
    import torch
    from mbpls_b200 import ingest
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(64, (20, 31, 12), 1, 3, seed=62)
    frames = {f"block{b}": pd.DataFrame(x, columns=[f"v{b}_{j}" for j in range(x.shape[1])]) for b, x in enumerate(X)}
    paths = []
    for name, df in frames.items():
        path = os.path.join(tmp_path, name + ".pkl")
        df.to_pickle(path)
        paths.append(path)
    blocks, names, columns = ingest.read_pickled_blocks(paths)
    assert names == list(frames) and all(b.is_pinned() and b.dtype == torch.float64 for b in blocks)
    assert list(columns[1]) == list(frames["block1"].columns)
    ref = _fit([x.copy() for x in X], Y.ravel().copy())
    for inputs in (blocks, list(frames.values()), ingest.pin_blocks(frames), ingest.pin_blocks(frames, dtype=torch.float32)):
        m = _fit(list(inputs), Y.ravel().copy())
        tol = 1e-5 if getattr(inputs[0], "dtype", None) == torch.float32 else 1e-12
        assert rel_err(m.beta_, ref.beta_) < tol
    with pytest.raises(FileNotFoundError):
        ingest.read_pickled_blocks([os.path.join(tmp_path, "missing.pkl")])
    # read-only sources are read, never copied or written
    ro = [x.copy() for x in X]
    for x in ro:
        x.setflags(write=False)
    assert rel_err(_fit(ro, Y.ravel().copy()).beta_, ref.beta_) < 1e-12
