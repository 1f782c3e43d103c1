"""Whole dense NIPALS fits in one kernel launch (csrc/smallfit.cu): the launch-latency regime of the reference.

The reference's README quickstart (README.rst:82-95) and the leave-one-out loops of its notebooks
(``cross_val_predict(MBPLS(n_components=k), X, y, cv=len(X))``, examples/real_world_applications/*.ipynb) fit matrices of
a few hundred kilobytes.  Through the streaming kernels such a fit is ~150 launches plus a host readback per component;
here one persistent CTA per fit runs the whole loop of mbpls/mbpls.py:821-983 on the device, and a grid of CTAs runs all
folds of a cross-validation at once.  This module packs the inputs (one host->device copy), launches, and unpacks the
results (one device->host copy).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from . import engine as E
from ._cabi import call
from .engine import F64, ptr, stream_ptr

# a fit runs in one CTA when its matrix has at most this many elements (1 MB of fp64): beyond it the single SM's L2
# bandwidth costs more than the launches it saves
SMALL_ELEMS = 1 << 17
# cross-validation: all folds at once when their private copies of the training data fit this many bytes
CV_WORKSPACE_BYTES = 4 << 30


def eligible(method: str, sparse: bool, group, n: int, p: int, force) -> bool:
    """force: the runtime option small_path (None = automatic by size; MBPLS_SMALL_PATH=0 in the environment turns the
    automatic choice off, which the test-suite uses to keep exercising the streaming kernels on small fixtures)."""
    if force is False or method != 'NIPALS' or sparse or group is not None or p < 1 or n < 1:
        return False
    if force is None and os.environ.get("MBPLS_SMALL_PATH", "1") == "0":
        return False
    return bool(force) or n * p <= SMALL_ELEMS


def pack_source(blocks: Sequence, Y, device):
    """Feature-major source matrix [X_1' ; ... ; X_B' ; Y'] ((p + q) x ldx, raw values) on the device in ONE copy, with the
    block offset table (B + 1 int32) behind it in the same buffer.  Host arrays are transposed into one pinned staging
    buffer; device tensors are concatenated on the device.  Returns (flat device buffer, n, sizes, p, q, ldx)."""
    n = int(blocks[0].shape[0])
    sizes = [int(b.shape[1]) for b in blocks]
    p, q = sum(sizes), int(Y.shape[1])
    ldx = (n + 1) // 2 * 2
    off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    tail = (len(off) + 1) // 2  # doubles that hold the int32 table
    if all(isinstance(b, torch.Tensor) and b.is_cuda for b in blocks):
        Yd = Y if isinstance(Y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(Y, dtype=np.float64))
        parts = [b.to(F64).t() for b in blocks] + [Yd.to(device=device, dtype=F64).t()]
        D = torch.zeros((p + q) * ldx + tail, dtype=F64, device=device)
        D[:(p + q) * ldx].view(p + q, ldx)[:, :n] = torch.cat(parts, dim=0)
        D[(p + q) * ldx:].view(torch.int32)[:len(off)] = torch.from_numpy(off).to(device)
        return D, n, sizes, p, q, ldx
    H = torch.empty((p + q) * ldx + tail, dtype=F64, pin_memory=True)
    Hn = H.numpy()
    M = Hn[:(p + q) * ldx].reshape(p + q, ldx)
    o = 0
    for b in blocks:
        a = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
        M[o:o + a.shape[1], :n] = a.T
        o += a.shape[1]
    Ya = Y.detach().cpu().numpy() if isinstance(Y, torch.Tensor) else np.asarray(Y)
    M[p:p + q, :n] = Ya.T
    if ldx > n:
        M[:, n:] = 0.0
    Hn[(p + q) * ldx:].view(np.int32)[:len(off)] = off
    return H.to(device, non_blocking=True), n, sizes, p, q, ldx


@dataclass
class Layout:
    """Offsets (in doubles) of the per-fit result sections inside the packed output buffer."""
    nfits: int
    p: int
    q: int
    B: int
    K: int
    ldw: int

    def __post_init__(self):
        p, q, B, K, ldw, F = self.p, self.q, self.B, self.K, self.ldw, self.nfits
        self.sizes = dict(stats=4 * p + 4 * q, Wt=K * p, W=K * p, P=K * p, Ts=K * ldw, U=K * ldw, Tb=B * K * ldw,
                          small=K * (q + 2 * B + 4) + B + 2, R=K * p, beta=q * p)
        self.off, o = {}, 0
        for name, sz in self.sizes.items():
            self.off[name] = o
            o += sz * F
        self.total = o

    def view(self, buf, name, f=0):
        sz = self.sizes[name]
        o = self.off[name] + f * sz
        return buf[o:o + sz]


def launch(D: torch.Tensor, n_src: int, p: int, q: int, ldx: int, block_off: Sequence[int], K: int, standardize: bool,
           norm_kind: int, max_tol: float, max_iter: int, train_sets: Optional[List[np.ndarray]],
           test_sets: Optional[List[np.ndarray]] = None):
    """Run ``len(train_sets)`` fits concurrently (train_sets None: one fit on all samples).  ``D`` is pack_source's buffer.
    Returns (layout, packed device result buffer, preds device tensor or None, buffers to keep alive)."""
    dev = D.device
    B = len(block_off) - 1
    boff_ptr = C.c_void_p(D.data_ptr() + 8 * (p + q) * ldx)  # the table pack_source put behind the matrix
    packed = None
    if train_sets is None:  # one fit on every sample of the source: no index tables at all
        F, ld_idx, ldw, ld_t = 1, n_src, (n_src + 1) // 2 * 2, 0
        iptr = lambda i: None
    else:
        F = len(train_sets)
        ld_idx = max(len(t) for t in train_sets)
        ldw = (ld_idx + 1) // 2 * 2
        tr = np.zeros((F, ld_idx), dtype=np.int32)
        cnt = np.zeros(F, dtype=np.int32)
        for f, t in enumerate(train_sets):
            tr[f, :len(t)] = t
            cnt[f] = len(t)
        ints = [tr.ravel(), cnt]
        ld_t = 0
        if test_sets is not None:
            ld_t = max(1, max(len(t) for t in test_sets))
            te = np.zeros((F, ld_t), dtype=np.int32)
            tcnt = np.zeros(F, dtype=np.int32)
            for f, t in enumerate(test_sets):
                te[f, :len(t)] = t
                tcnt[f] = len(t)
            ints += [te.ravel(), tcnt]
        packed = torch.from_numpy(np.concatenate(ints)).to(dev, non_blocking=True)  # every index table in one copy
        offs = np.cumsum([0] + [len(a) for a in ints])
        iptr = lambda i: C.c_void_p(packed.data_ptr() + 4 * int(offs[i]))
    lay = Layout(F, p, q, B, K, ldw)
    out = torch.empty(lay.total, dtype=F64, device=dev)
    work = torch.empty(F * (p + q) * ldw, dtype=F64, device=dev)
    sstride = call("mbpls_smallfit_scratch_doubles", p, B, ldw)
    scratch = torch.empty(F * sstride, dtype=F64, device=dev)
    preds = torch.full((K, n_src, q), float("nan"), dtype=F64, device=dev) if test_sets is not None else None
    base = out.data_ptr()
    sec = lambda name: C.c_void_p(base + 8 * lay.off[name])
    args = _cabi.SmallFitArgs(
        n_src=n_src, p=p, B=B, q=q, K=K, nfits=F, ldx=ldx, Xsrc=D.data_ptr(), Ysrc=D.data_ptr() + 8 * p * ldx, block_off=boff_ptr,
        standardize=1 if standardize else 0, norm_kind=norm_kind, max_iter=int(min(max_iter, 2**31 - 1)), max_tol=float(max_tol),
        train_idx=iptr(0), train_cnt=iptr(1), ld_idx=ld_idx,
        test_idx=iptr(2) if test_sets is not None else None, test_cnt=iptr(3) if test_sets is not None else None, ld_tidx=ld_t,
        ldw=ldw, Xw=work.data_ptr(), Yw=work.data_ptr() + 8 * F * p * ldw, stats=sec("stats"), Wt=sec("Wt"), W=sec("W"), P=sec("P"),
        Ts=sec("Ts"), U=sec("U"), Tb=sec("Tb"), small=sec("small"), R=sec("R"), beta=sec("beta"),
        preds=preds.data_ptr() if preds is not None else None, scratch=scratch.data_ptr(), scratch_stride=sstride)
    call("mbpls_smallfit_nipals_f64", C.byref(args), stream_ptr(dev))
    return lay, out, preds, (packed, work, scratch)


def unpack_small(small: np.ndarray, K: int, q: int, B: int):
    """Sections of the per-fit `small` block (see include/mbpls_b200.h)."""
    o = 0
    def take(sz):
        nonlocal o
        v = small[o:o + sz]
        o += sz
        return v
    return dict(V=take(K * q).reshape(K, q), A=take(K * B).reshape(K, B), pssb=take(K * B).reshape(K, B), tt=take(K), vv=take(K),
                diff=take(K), trips=take(K), varxb=take(B), vary=float(take(1)[0]), singular=bool(take(1)[0] != 0.0))
