"""Device-resident cross-validation driver.

Every real-world notebook of the reference runs
``sklearn.model_selection.cross_val_predict(MBPLS(n_components=k), X, y, cv=len(X))`` for a range of ``k``
(``examples/real_world_applications/*.ipynb``): n refits per k, each of which re-validates, copies and -- with a
GPU estimator -- would re-upload X.  Here X and Y are uploaded once (feature-major); each fold gathers its training
samples on the device, runs the normal fit kernels on them, and predicts its held-out samples; with
``n_components_list`` the predictions for every smaller model are read off the *same* fit, because a K-component
NIPALS / UNIPALS / SIMPLS / KERNEL fit contains all of its prefixes (``beta_k = R_k V_k'`` with
``R_k = W_k pinv(P_k' W_k)``, mbpls/mbpls.py:986-989).

The result equals ``cross_val_predict(clone(estimator), X, y, cv=cv)`` (tests/test_gpu_cv.py).
"""
from __future__ import annotations

import warnings
from typing import Iterable, Optional, Sequence

import numpy as np
import torch
from sklearn.base import clone

from . import engine as E
from .engine import F64
from .mbpls import MBPLS, _as_2d_source, _is_block_list

__all__ = ["cross_val_predict"]


def _folds(n: int, cv) -> list:
    if isinstance(cv, (int, np.integer)):
        from sklearn.model_selection import KFold
        return [(tr, te) for tr, te in KFold(n_splits=int(cv)).split(np.arange(n))]
    if hasattr(cv, "split"):
        return [(tr, te) for tr, te in cv.split(np.arange(n))]
    return [(np.asarray(tr), np.asarray(te)) for tr, te in cv]


def _prefix_betas(model: MBPLS, ks: Sequence[int], device) -> list:
    """beta for the first k components, k in ks, from one fitted model (device tensors q x p each)."""
    dev = model._device_model(device)
    shard = dev["shard"]
    p = shard.p_local
    group, _, _ = model._group_info()
    P, V = dev["P"], dev["V"]
    if model.method == 'SIMPLS':
        # SIMPLS: R and Q columns do not depend on later components (mbpls.py:1004-1013, :1044)
        return [E.right_multiply(dev["R"][:k], p, None, V[:k].contiguous()) for k in ks]
    Wfull = model.__dict__.get("_cv_weights")  # K x p un-normalised / concatenated weights kept by fit for CV
    out = []
    for k in ks:
        Wk, Pk = Wfull[:k], P[:k]
        if model.method == 'NIPALS':
            coln = torch.sqrt(E.rows_sumsq(Wk, p, group))
            M = E.small_pinv(E.gram(Pk, Wk, p, group) / coln.view(1, -1))
            R = E.right_multiply(Wk, p, 1.0 / coln, M)
        else:
            M = E.small_pinv(E.gram(Pk, Wk, p, group))
            R = E.right_multiply(Wk, p, None, M)
        out.append(E.right_multiply(R, p, None, V[:k].contiguous()))
    return out


def _batched_small_cv(estimator, blocks, Ysrc, n, q, sizes, shard, K, ks, folds, device):
    """All folds of a dense NIPALS cross-validation in ONE kernel launch (csrc/smallfit.cu): one CTA per fold gathers its
    training samples from the shared source, fits K components with the loop on the device, and predicts its held-out samples
    for every prefix 1..K of the model.  Returns {k: (n, q) array}, or None when the problem is not eligible (-> fold loop)."""
    from . import smallfit as SF
    rt = estimator._runtime()
    p = sum(sizes)
    force = rt["small_path"]
    if force is False or estimator.method != 'NIPALS' or estimator.sparse_data or not folds:
        return None
    ntr_max = max(len(tr) for tr, _ in folds)
    if force is None:
        import os
        if os.environ.get("MBPLS_SMALL_PATH", "1") == "0":
            return None
        if ntr_max * p > 8 * SF.SMALL_ELEMS or len(folds) * (p + q) * ntr_max * 8 > SF.CV_WORKSPACE_BYTES:
            return None
    from .mbpls import _check_limits
    _check_limits(K, q, len(sizes))
    for a in blocks + [Ysrc]:
        ok = bool(torch.isfinite(a).all()) if isinstance(a, torch.Tensor) else bool(np.isfinite(a).all())
        if not ok:
            raise ValueError("Input contains NaN or infinity.")
    D, n, sizes, p, q, ldx = SF.pack_source(blocks, Ysrc, device)
    lay, out, preds_d, keep = SF.launch(D, n, p, q, ldx, shard.block_off, K, bool(estimator.standardize),
                                        E.norm_kind_of(estimator.nipals_convergence_norm), estimator.max_tol, rt["max_iter"],
                                        [np.asarray(tr, dtype=np.int32) for tr, _ in folds],
                                        [np.asarray(te, dtype=np.int32) for _, te in folds])
    sm_all = E.to_host(out[lay.off["small"]:lay.off["small"] + lay.sizes["small"] * len(folds)]).reshape(len(folds), -1)
    if np.any(sm_all[:, -1] != 0.0):  # some fold has more components than rank: the fold loop applies the pseudo-inverse
        return None
    Pm = E.to_host(preds_d)  # K x n x q
    return {k: np.ascontiguousarray(Pm[k - 1]) for k in (ks or [K])}


def cross_val_predict(estimator: MBPLS, X, y, cv=5, n_components_list: Optional[Iterable[int]] = None):
    """Out-of-fold predictions of ``estimator`` (an unfitted ``mbpls_b200.MBPLS``) with the data kept on the GPU.

    Returns an (n, q) array like ``sklearn.model_selection.cross_val_predict``; with ``n_components_list`` a dict
    ``{k: (n, q) array}`` obtained from one fit per fold with ``max(k)`` components.
    With a process group in the estimator's runtime options the FOLDS are spread over the ranks (rank r fits folds r, r + world,
    ...; the data is replicated, every fit runs on one GPU) and the out-of-fold predictions are summed across the ranks, so every
    rank returns the complete result.
    """
    rt = estimator._runtime()
    group = rt["group"]
    device = E.require_cuda(rt["device"])
    blocks = X if _is_block_list(X) else [X]
    blocks = [_as_2d_source(b, "X") for b in blocks]
    Ysrc = y if isinstance(y, torch.Tensor) else np.asarray(y)
    y1d = Ysrc.ndim == 1
    if y1d:
        Ysrc = Ysrc.reshape(-1, 1)
    Ysrc = _as_2d_source(Ysrc, "y")
    n, q = int(Ysrc.shape[0]), int(Ysrc.shape[1])
    sizes = [int(b.shape[1]) for b in blocks]
    ks = None if n_components_list is None else sorted(set(int(k) for k in n_components_list))
    K = estimator.n_components if ks is None else max(ks)
    folds = _folds(n, cv)
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        folds = folds[rank::world]

    def finish(preds):
        """dict k -> (n, q) with NaN outside this rank's folds -> the complete arrays on every rank"""
        if group is not None:
            keys = sorted(preds)
            stack = torch.from_numpy(np.stack([preds[k] for k in keys])).to(device)
            seen = (~torch.isnan(stack)).to(F64)
            stack = torch.nan_to_num(stack, nan=0.0)
            dist.all_reduce(stack, group=group)
            dist.all_reduce(seen, group=group)
            stack = torch.where(seen > 0, stack / seen.clamp(min=1.0), torch.full_like(stack, float("nan")))
            host = stack.cpu().numpy()
            preds = {k: host[i] for i, k in enumerate(keys)}
        if ks is None:
            return preds[K].ravel() if y1d else preds[K]
        return {k: (v.ravel() if y1d else v) for k, v in preds.items()}

    with torch.cuda.device(device):
        shard = E.ShardMap.build(sizes, 0, 1)
        batched = _batched_small_cv(estimator, blocks, Ysrc, n, q, sizes, shard, K, ks, folds, device) if folds else None
        if batched is not None:
            return finish(batched)
        Xraw = E.ingest_blocks(blocks, n, shard, device)          # p x ld, uploaded once
        Yraw = E.alloc_feature_major(q, n, device)
        E.ingest_feature_major(Ysrc, n, 0, q, Yraw, device)
        off = shard.block_off
        preds = {k: np.full((n, q), np.nan) for k in (ks or [K])}
        for tr, te in folds:
            tr_d = torch.as_tensor(np.asarray(tr), device=device, dtype=torch.long)
            te_d = torch.as_tensor(np.asarray(te), device=device, dtype=torch.long)
            ntr, nte = len(tr), len(te)
            Xtr = E.alloc_feature_major(shard.p_local, ntr, device)
            Xtr[:, :ntr] = Xraw.index_select(1, tr_d)              # gather of the training samples (device copy)
            Ytr = Yraw.index_select(1, tr_d).t().contiguous()      # ntr x q
            Xte = E.alloc_feature_major(shard.p_local, nte, device)
            Xte[:, :nte] = Xraw.index_select(1, te_d)
            m = clone(estimator)
            m.set_params(n_components=K, copy=False)
            m.set_runtime(device=device, materialize=False)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m.fit([Xtr[off[b]:off[b + 1], :ntr].t() for b in range(len(sizes))], Ytr)
                te_blocks = [Xte[off[b]:off[b + 1], :nte].t() for b in range(len(sizes))]
                if ks is None:
                    preds[K][te] = m.predict(te_blocks)
                else:
                    dev, sh, Xt_new, mm, mean, scale = m._prepare_new_X(te_blocks, device, scaled_copy=False)
                    for k, beta_k in zip(ks, _prefix_betas(m, ks, device)):
                        Yh = E.skinny_gemm(Xt_new, mm, beta_k, sh.block_off, None, mean, scale)
                        if m.standardize:
                            _, _, ymean, yscale = m._device_scalers(sh, device)
                            E.call("mbpls_scaler_inverse_f64", E.ptr(Yh), Yh.shape[1], mm, q, E.ptr(ymean), E.ptr(yscale),
                                   E.stream_ptr(device))
                        preds[k][te] = Yh[:, :mm].cpu().numpy().T
    return finish(preds)
