"""KERNEL / UNIPALS / SIMPLS fit bodies (mbpls/mbpls.py:384-807, :995-1048) -- filled in below."""
from __future__ import annotations


def fit(model, Xt, Yt, n, q, shard, boff_dev, zss, group, device):
    raise NotImplementedError(f"method {model.method!r} is not implemented yet in mbpls_b200")
