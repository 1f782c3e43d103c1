"""SIMPLS / UNIPALS / KERNEL fit bodies (reference: mbpls/mbpls.py:995-1048, :384-574, :576-807) on the
feature-major device matrix.

Algebra that replaces the reference's dense p x p / n x n work:
* ``S = X'YY'X = C C'`` with ``C = X'Y`` (p x q): its top left singular vector is ``C c / |C c|`` with ``c`` the top
  eigenvector of the q x q matrix ``C'C``  (:396-400, :584-592).
* ``S = XX'YY' = A Y'`` with ``A = X (X'Y)`` (n x q): the top left singular vector of ``A Y'`` is ``A c / |A c|`` with
  ``c = L z``, ``Y'Y = L L'``, ``z`` the top eigenvector of ``L' (A'A) L``  (:489-493, :704-715).
* KERNEL sandwich deflations (:630-633, :722-725) are rank-1 / rank-2 updates because ``w'VAR = den * p`` and
  ``D = I - ts ts'`` is symmetric.
The q x q / K x K eigen-problems and pseudo-inverses run on the device too (one-sided Jacobi SVD, csrc/smalllin.cu),
so a fit enqueues asynchronously without per-component host round-trips.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
import weakref

import numpy as np
import torch

from . import engine as E
from .engine import F64, call, ptr, stream_ptr


# ------------------------------------------------------------------------------------------------
# thin wrappers
# ------------------------------------------------------------------------------------------------
def xt_multi(Xt, n, M):
    """(C x ld) right-hand sides -> C x p_local with out[c][j] = x_j . M[c]."""
    p, ld = Xt.shape
    Cc = M.shape[0]
    out = torch.zeros((Cc, max(p, 1)), dtype=F64, device=Xt.device)
    if p > 0:
        chunks = call("mbpls_xt_multi_chunks", n, p)
        if chunks > 1:  # few, long features: cut the sample axis, add the partial products in chunk order
            part = torch.zeros((chunks, Cc * out.stride(0)), dtype=F64, device=Xt.device)
            call("mbpls_xt_multi_split_f64", ptr(Xt), ld, n, p, ptr(M), M.stride(0), Cc, ptr(part), out.stride(0), chunks,
                 stream_ptr(Xt.device))
            call("mbpls_reduce_chunks_f64", ptr(part), chunks, Cc * out.stride(0), ptr(out), stream_ptr(Xt.device))
        else:
            call("mbpls_xt_multi_f64", ptr(Xt), ld, n, p, ptr(M), M.stride(0), Cc, ptr(out), out.stride(0), stream_ptr(Xt.device))
    return out[:, :p]


def xt_vec(Xt, n, u, boff_dev, B, uu=None):
    """x_j . u (/ uu) for every local feature -> p_local vector (the NIPALS xtu kernel)."""
    p, ld = Xt.shape
    w = torch.zeros(max(p, 1), dtype=F64, device=Xt.device)
    call("mbpls_nipals_xtu_f64", ptr(Xt), ld, n, p, ptr(u), ptr(uu), ptr(boff_dev), B, ptr(w), None, 0, None,
         stream_ptr(Xt.device))
    return w[:p]


def lincomb_sub(base, V, K, coef, length):
    out = base.clone()
    call("mbpls_lincomb_sub_f64", ptr(out), ptr(base), ptr(V), V.stride(0) if K > 0 else 0, K, ptr(coef), length,
         stream_ptr(base.device))
    return out


def normalize_(t, n, center=False):
    if n >= (1 << 16) and not center:
        # long vectors: the single-CTA kernel below makes three passes through one SM (435 us for 1 M samples); sum of squares
        # + a grid-wide scaling pass instead
        nrm = torch.sqrt(E.rows_sumsq(t.view(1, -1), n))
        scale_rows_(t.view(1, -1), n, nrm, True)
        return nrm
    nrm = torch.zeros(1, dtype=F64, device=t.device)
    call("mbpls_center_normalize_f64", ptr(t), n, 1 if center else 0, 1, ptr(nrm), stream_ptr(t.device))
    return nrm


def scale_rows_(M, n, scale, divide):
    rows = M.shape[0] if M.dim() == 2 else 1
    call("mbpls_rows_scale_f64", ptr(M), M.stride(0) if M.dim() == 2 else n, rows, n, ptr(scale), 1 if divide else 0,
         stream_ptr(M.device))


def normalize_over_features_(w, p, group):
    """w <- w / |w| with the norm taken over the *global* feature axis."""
    nrm = torch.sqrt(E.rows_sumsq(w.view(1, -1), p, group))
    if p > 0:
        scale_rows_(w.view(1, -1), p, nrm, True)
    return nrm


def block_sumsq(w, boff_dev, B, group):
    out = torch.zeros(B, dtype=F64, device=w.device)
    call("mbpls_block_sumsq_f64", ptr(w), ptr(boff_dev), B, ptr(out), stream_ptr(w.device))
    E.allreduce_(out, group)
    return out


def scale_by_block(w, boff_dev, B, a, p):
    out = torch.empty_like(w)
    call("mbpls_scale_by_block_f64", ptr(w), ptr(boff_dev), B, ptr(a), ptr(out), p, stream_ptr(w.device))
    return out


def rank1_update_(Xt, n, ts, pvec):
    p, ld = Xt.shape
    call("mbpls_rank1_update_f64", ptr(Xt), ld, n, p, ptr(ts), ptr(pvec), stream_ptr(Xt.device))


class BlockProducts:
    """red[b][i] = sum_{j in block b} w[j] x_ij for all blocks in one pass (the NIPALS xw + reduce kernels)."""

    def __init__(self, Xt, n, block_off, group):
        dev = Xt.device
        self.Xt, self.n, self.group = Xt, n, group
        self.B = len(block_off) - 1
        f0, f1, bso = E.make_splits(block_off, n, E.sm_count(dev))
        self.ns = len(f0)
        self.sf0, self.sf1, self.sbso = E._i32(f0, dev), E._i32(f1, dev), E._i32(bso, dev)
        ld = Xt.shape[1]
        self.Tnum = torch.zeros((max(self.ns, 1), ld), dtype=F64, device=dev)
        self.dummy = torch.zeros(max(self.B, 1), dtype=F64, device=dev)

    def __call__(self, w):
        Xt, n, B = self.Xt, self.n, self.B
        ld = Xt.shape[1]
        st = stream_ptr(Xt.device)
        red = torch.zeros(B * ld + B, dtype=F64, device=Xt.device)
        call("mbpls_nipals_xw_f64", ptr(Xt), ld, n, ptr(w), ptr(self.sf0), ptr(self.sf1), self.ns, ptr(self.Tnum), None, ld, 0,
             None, st)
        call("mbpls_nipals_reduce_partials_f64", ptr(self.Tnum), None, ld, n, B, ptr(self.sbso), ptr(self.dummy), 0, ptr(red),
             0, None, st)
        E.allreduce_(red, self.group)
        return red[:B * ld].view(B, ld)


def sum_rows(M, length):
    """out[i] = sum_b M[b][i] (fixed order)."""
    out = torch.zeros(M.shape[1], dtype=F64, device=M.device)
    Mc = M.contiguous()
    call("mbpls_reduce_chunks_f64", ptr(Mc), Mc.shape[0], Mc.shape[1], ptr(out), stream_ptr(M.device))
    return out


def _bip_corrected(a, sizes):
    a = np.asarray(a, dtype=np.float64).ravel()
    if a.size == 1:
        return np.array([1.0])
    sizes = np.asarray(sizes, dtype=np.float64)
    corrected = a * (1.0 - sizes / sizes.sum())
    return corrected / corrected.sum()


def _finish_common(model, shard, n, q, R, beta, P, Ts, U, V, Wb, Tb, lazy_extra=None):
    """Register lazily materialised attributes shared by the three methods."""
    mw = weakref.proxy(model)  # the closures below must not keep the model alive (model -> _lazy -> closure -> model)
    B = len(shard.sizes)
    lazy = {
        "Ts_": lambda: mw._gather_samples(Ts, n),
        "U_": lambda: mw._gather_samples(U, n) if U is not None else np.empty((n, 0)),
        "V_": lambda: E.to_host(V, transpose=True),
        "P_": lambda: mw._features_T(P, shard, True),
        "R_": lambda: mw._features_T(R, shard, False),
        "beta_": lambda: mw._features_T(beta, shard, False),
    }
    if Wb is not None:
        lazy["W_"] = lambda: mw._features_T(Wb, shard, True)
    if Tb is not None:
        lazy["T_"] = lambda: [mw._gather_samples(Tb[b], n) for b in range(B)]
    if lazy_extra:
        lazy.update(lazy_extra)
    model.__dict__["_lazy"] = lazy
    model.__dict__["_dev"] = dict(shard=shard, R=R, beta=beta, P=P, V=V, **({"W": Wb} if Wb is not None else {}))


# ------------------------------------------------------------------------------------------------
def fit(model, Xt, Yt, n, q, shard, boff_dev, zss, group, device):
    method = model.method
    if method == 'SIMPLS':
        return _fit_simpls(model, Xt, Yt, n, q, shard, boff_dev, group, device)
    if method == 'UNIPALS':
        return _fit_unipals(model, Xt, Yt, n, q, shard, boff_dev, zss, group, device)
    if method == 'KERNEL':
        return _fit_kernel(model, Xt, Yt, n, q, shard, boff_dev, zss, group, device)
    raise NameError('Method you called is unknown')


# ---- SIMPLS (mbpls.py:995-1048) ------------------------------------------------------------------
def _fit_simpls(model, Xt, Yt, n, q, shard, boff_dev, group, device):
    mw = weakref.proxy(model)
    warnings.warn("Method 'SIMPLS' does not calculate A_ and T_!")  # :996
    K, B = int(model.n_components), len(shard.sizes)
    p, ld = Xt.shape
    St = xt_multi(Xt, n, Yt).contiguous()  # q x p : S = X'Y (:998)
    if p == 0:
        St = torch.zeros((q, 1), dtype=F64, device=device)[:, :0]
    R, Tm, P, Q, U, V, W = (torch.zeros((K, max(p, 1)), dtype=F64, device=device), torch.zeros((K, ld), dtype=F64, device=device),
                           torch.zeros((K, max(p, 1)), dtype=F64, device=device), torch.zeros((K, q), dtype=F64, device=device),
                           torch.zeros((K, ld), dtype=F64, device=device), torch.zeros((K, max(p, 1)), dtype=F64, device=device),
                           torch.zeros((K, max(p, 1)), dtype=F64, device=device))
    ldS = St.stride(0) if p > 0 else 1
    for k in range(K):
        qv = E.small_top_eigvec(E.gram(St, St, p, group))  # top singular vector of S'S, q x q (:1001)
        r = E.right_multiply(St, p, None, qv.view(q, 1))[0].contiguous()  # r = S q (:1004)
        W[k, :p] = r
        t = E.skinny_gemm(Xt, n, r.view(1, -1), shard.block_off, group, dense=True)[0].contiguous()  # t = X r (:1006)
        normt = normalize_(t, n, center=True)  # :1007-1009
        if p > 0:
            scale_rows_(r.view(1, -1), p, normt, True)  # :1010
        pv = xt_vec(Xt, n, t, boff_dev, B)  # p = X't (:1011)
        qk = E.gram(Yt, t.view(1, -1), n)[:, 0].contiguous()  # q = Y't (:1012)
        u = E.skinny_gemm(Yt, n, qk.view(1, -1), [0, q], dense=True)[0].contiguous()  # u = Y q (:1013)
        v = pv.clone() if p > 0 else pv
        if k > 0:  # :1015-1017
            cv = E.gram(V[:k, :p], pv.view(1, -1), p, group)[:, 0].contiguous()
            if p > 0:
                v = lincomb_sub(pv.contiguous(), V, k, cv, p)
            cu = E.gram(Tm[:k], u.view(1, -1), n)[:, 0].contiguous()
            u = lincomb_sub(u, Tm, k, cu, n)
        normalize_over_features_(v, p, group)  # :1018
        vS = E.gram(St, v.view(1, -1), p, group)[:, 0].contiguous()  # v'S (q)
        if p > 0:
            call("mbpls_rank1_update_f64", ptr(St), ldS, p, q, ptr(v), ptr(vS), stream_ptr(device))  # S -= v v'S (:1019)
        normalize_(u, n)
        R[k, :p], Tm[k], P[k, :p], Q[k], U[k], V[k, :p] = r, t, pv, qk, u, v
    R, P, W = R[:, :p], P[:, :p], W[:, :p]
    beta = E.right_multiply(R, p, None, Q.contiguous())  # beta = R Q' (:1044)
    model.explained_var_x_, model.explained_var_y_ = [], []
    model.explained_var_xblocks_ = np.empty((B, 0))
    model.W_non_normal_ = [np.empty((s, 0)) for s in shard.sizes]
    model.W_concat_ = np.empty((shard.p_global, 0))
    _finish_common(model, shard, n, q, R, beta, P, Tm, U, Q, None, None,
                   {"W_": lambda: mw._features_T(W, shard, False)})


# ---- UNIPALS (mbpls.py:384-574) ------------------------------------------------------------------
def _gram_route_pays(n_global, p_global, K):
    """UNIPALS with n >= p through the Gram matrices: one SYRK (n p^2 executed flops at ~30 TFLOP/s) against K components of
    at least three passes over X at ~5 TB/s; also the p x p matrix (plus split-K partials) must stay small."""
    if n_global < p_global or p_global < 1:
        return False
    syrk_s = float(n_global) * p_global * p_global / 3.0e13
    stream_s = K * 3.0 * 8.0 * n_global * p_global / 5.0e12
    return p_global <= 16384 and syrk_s < stream_s


def _fit_unipals_gram(model, Xt, Yt, n, q, shard, boff_dev, zss, device, rows_group=None, n_global=None):
    """UNIPALS for n >= p without a single per-component pass over X.

    The reference rebuilds S = X'YY'X from the deflated X in every component and then makes four more passes over X
    (t_b = X_b w_b, ts = X w, p = X'ts, X -= ts p'; :396-443).  With n >> p all of that lives in p-space: keep VAR = X'X and
    C = X'Y (one tensor-core SYRK and one pass, the KERNEL method's cross-products, :586-587) and deflate THEM,
        ts = X w / |X w|,  |X w|^2 = w'VAR w,  p = X'ts = VAR w / |X w|,  v = Y'ts = C'w / |X w|,
        X <- X - ts p'   ==>   VAR <- VAR - p p',  C <- C - p v'      (ts'ts = 1, X'ts = p),
    then recover the n-space outputs after the loop: Ts = X R (R = W pinv(P'W), unit columns), block scores
    T_b = X_b W_b - Ts striu(P_b'W_b) (the K sequential deflations collapse to a triangular correction), U = Y v / v'v.
    Same quantities as the streaming form up to rounding (4e-12 against the numpy reference over 30 components); C5 (n = 1 M,
    p = 2,000, K = 30): 6 reads of X + one SYRK instead of ~150 passes.  rows_group: sample axis sharded over the ranks
    (VAR, C and the few squared norms are all-reduced once; everything else is row-local or replicated)."""
    K, B = int(model.n_components), len(shard.sizes)
    p, ld = Xt.shape
    rg = rows_group
    st = stream_ptr(device)
    if zss is None:
        zss = E.feature_sumsq(Xt, n)
    varxb = model._block_sums(zss, boff_dev, B, rg)
    vary_t = E.segsum(E.feature_sumsq(Yt, n), E._i32([0, q], device), 1).clone()
    E.allreduce_(vary_t, rg)
    vary = float(vary_t.item())
    Ct = xt_multi(Xt, n, Yt).contiguous()  # C' = (X'Y)' : q x p
    ldv = (p + 15) // 16 * 16
    VAR = crossprod(Xt, Xt, p, p, n, True, ldv)  # X'X on the FP64 tensor cores
    E.allreduce_(Ct, rg)
    E.allreduce_(VAR, rg)
    Wc = torch.zeros((K, p), dtype=F64, device=device)
    Wb = torch.zeros((K, p), dtype=F64, device=device)
    P = torch.zeros((K, p), dtype=F64, device=device)
    V = torch.zeros((K, q), dtype=F64, device=device)
    U = torch.zeros((K, ld), dtype=F64, device=device)
    A_dev = torch.zeros((K, B), dtype=F64, device=device)
    pss_dev = torch.zeros((K, B), dtype=F64, device=device)
    vv_dev = torch.zeros(K, dtype=F64, device=device)
    for k in range(K):
        c = E.small_top_eigvec(E.gram(Ct, Ct, p))  # top eigenvector of C'C (q x q) -> w = C c / |C c| (:396-402)
        w = E.right_multiply(Ct, p, None, c.view(q, 1))[0].contiguous()
        normalize_over_features_(w, p, None)
        Vw = torch.zeros(p, dtype=F64, device=device)
        call("mbpls_dense_gemv_f64", ptr(VAR), ldv, p, p, ptr(w), ptr(Vw), st)
        tn = torch.sqrt(E.gram(w.view(1, -1), Vw.view(1, -1), p).view(-1))  # |X w|
        v = E.gram(Ct, w.view(1, -1), p)[:, 0].contiguous()
        scale_rows_(v.view(1, -1), q, tn, True)   # v = Y'ts / ts'ts (:420)
        pv = Vw
        scale_rows_(pv.view(1, -1), p, tn, True)  # p = X'ts / ts'ts (:427)
        a = block_sumsq(w, boff_dev, B, None)     # :405-408
        Wb[k] = scale_by_block(w, boff_dev, B, a, p)
        A_dev[k] = a
        u = E.skinny_gemm(Yt, n, v.view(1, -1), [0, q], dense=True)[0].contiguous()  # u = Y v / v'v, unit length (:423-424)
        if rg is None:
            normalize_(u, n)
        else:
            scale_rows_(u.view(1, -1), n, torch.sqrt(E.rows_sumsq(u.view(1, -1), n, rg)), True)
        pss_dev[k] = block_sumsq(pv, boff_dev, B, None)
        vv_dev[k:k + 1] = E.rows_sumsq(v.view(1, -1), q)
        # X <- X - ts p' (:443), in p-space
        call("mbpls_rank1_update_f64", ptr(Ct), Ct.stride(0), p, q, ptr(pv), ptr(v), st)
        call("mbpls_dense_rank2_f64", ptr(VAR), ldv, p, p, ptr(pv), ptr(pv), None, 0.0, 0.0, -1.0, -1, -1, st)
        Wc[k], P[k], V[k], U[k] = w, pv, v, u
    del VAR
    M = E.small_pinv(E.gram(P, Wc, p))
    R = E.right_multiply(Wc, p, None, M)  # :476
    beta = E.right_multiply(R, p, None, V.contiguous())  # :477
    Ts = E.skinny_gemm(Xt, n, R, shard.block_off, dense=True)  # ts_k = X_k w_k / |X_k w_k| = X r_k: unit columns by construction
    Tb = torch.zeros((B, K, ld), dtype=F64, device=device)
    for b in range(B):
        o0, o1 = shard.block_off[b], shard.block_off[b + 1]
        full = E.skinny_gemm(Xt[o0:o1], n, Wb[:, o0:o1], [0, o1 - o0], dense=True)  # X_b W_b : K x ld
        Cb = torch.triu(E.gram(P[:, o0:o1], Wb[:, o0:o1], o1 - o0), diagonal=1).contiguous()  # striu(P_b'W_b)
        Tb[b, 0] = full[0]
        for k in range(1, K):  # t_b,k = X_b^(k) w_b,k = X_b w_b,k - sum_{j<k} ts_j (p_b,j . w_b,k)  (:410-413 after :443)
            Tb[b, k] = lincomb_sub(full[k].contiguous(), Ts, k, Cb[:k, k].contiguous(), n)
    A = A_dev.cpu().numpy().T.copy()
    pssb_h, vv_h = pss_dev.cpu().numpy(), vv_dev.cpu().numpy()
    evx = [float(pssb_h[k].sum() / varxb.sum()) for k in range(K)]  # ((ts p')**2).sum() = ts'ts p'p with ts'ts = 1 (:429-448)
    evy = [float(vv_h[k] / vary) for k in range(K)]
    evxb = (pssb_h / varxb[None, :]).T.copy()
    model.A_ = A
    model.A_corrected_ = np.stack([_bip_corrected(A[:, k], shard.sizes) for k in range(K)], axis=1)
    model.explained_var_x_, model.explained_var_y_, model.explained_var_xblocks_ = evx, evy, evxb
    model.W_non_normal_ = [np.empty((s_, 0)) for s_ in shard.sizes]
    model.W_concat_ = np.empty((shard.p_global, 0))
    _finish_common(model, shard, n, q, R, beta, P, Ts, U, V, Wb, Tb)
    model.__dict__["_cv_weights"] = Wc


def _fit_unipals(model, Xt, Yt, n, q, shard, boff_dev, zss, group, device, rows_group=None, n_global=None):
    """group: the FEATURE axis is sharded over it (contractions over p are all-reduced).  rows_group: the SAMPLE axis is
    sharded instead (n >= p): Xt / Yt hold this rank's n rows of all features; per component the sums
    over samples -- X'Y (p x q), Y'ts (q), X'ts (p) and the two squared norms -- are all-reduced, everything indexed by
    features is replicated and everything indexed by samples (block scores, ts, u, the deflation) stays row-local."""
    K, B = int(model.n_components), len(shard.sizes)
    p, ld = Xt.shape
    pg = shard.p_global
    rg = rows_group
    ng = n if n_global is None else n_global
    if rg is not None and (group is not None or ng < pg):
        raise NotImplementedError("row-sharded UNIPALS covers n >= p with a replicated feature axis")
    route = model._runtime().get("unipals_route")
    if group is None and ng >= pg and p > 0 and route != "stream" and (route == "gram" or _gram_route_pays(ng, pg, K)):
        return _fit_unipals_gram(model, Xt, Yt, n, q, shard, boff_dev, zss, device, rows_group=rg, n_global=n_global)

    def normalize_rows_(t):  # unit norm over the *global* sample axis
        if rg is None:
            return normalize_(t, n)
        nrm = torch.sqrt(E.rows_sumsq(t.view(1, -1), n, rg))
        scale_rows_(t.view(1, -1), n, nrm, True)
        return nrm

    blockprod = BlockProducts(Xt, n, shard.block_off, group)
    if zss is None:
        zss = E.feature_sumsq(Xt, n)
    varxb = model._block_sums(zss, boff_dev, B, group if rg is None else rg)
    vary_t = E.segsum(E.feature_sumsq(Yt, n), E._i32([0, q], device), 1).clone()
    E.allreduce_(vary_t, rg)
    vary = float(vary_t.item())
    Wc = torch.zeros((K, max(p, 1)), dtype=F64, device=device)   # eigenv columns ("weights")
    Wb = torch.zeros((K, max(p, 1)), dtype=F64, device=device)   # block-normalised
    P = torch.zeros((K, max(p, 1)), dtype=F64, device=device)
    Ts, U = torch.zeros((K, ld), dtype=F64, device=device), torch.zeros((K, ld), dtype=F64, device=device)
    V = torch.zeros((K, q), dtype=F64, device=device)
    Tb = torch.zeros((B, K, ld), dtype=F64, device=device)
    evx, evy, evxb = [], [], np.zeros((B, K))
    GY = E.gram(Yt, Yt, n)
    A_dev = torch.zeros((K, B), dtype=F64, device=device)
    pss_dev = torch.zeros((K, B), dtype=F64, device=device)
    tt_dev = torch.zeros(K, dtype=F64, device=device)
    vv_dev = torch.zeros(K, dtype=F64, device=device)
    for k in range(K):
        Ct = xt_multi(Xt, n, Yt).contiguous()  # (X'Y)' from the *deflated* X (:396 / :489)
        E.allreduce_(Ct, rg)
        if ng >= pg:  # :388-424
            c = E.small_top_eigvec(E.gram(Ct, Ct, p, group))
            w = E.right_multiply(Ct, p, None, c.view(q, 1))[0].contiguous()
            normalize_over_features_(w, p, group)
            raw = blockprod(w)  # X_b w (un-normalised block part)
            ts = sum_rows(raw, ld)  # X w (:416)
            normalize_rows_(ts)
            tt = E.rows_sumsq(ts.view(1, -1), n, rg)
            v = E.gram(Yt, ts.view(1, -1), n, rg)[:, 0].contiguous()
            scale_rows_(v.view(1, -1), q, tt, True)  # v = Y'ts / ts'ts (:420)
            u = E.skinny_gemm(Yt, n, v.view(1, -1), [0, q], dense=True)[0].contiguous()  # :423-424
            normalize_rows_(u)
        else:  # :481-517
            At = E.skinny_gemm(Xt, n, Ct, shard.block_off, group, dense=True)  # (X X'Y)' : q x ld
            c = E.small_top_sv_product(GY, E.gram(At, At, n))
            ts = E.right_multiply(At, ld, None, c.view(q, 1))[0].contiguous()
            normalize_(ts, n)
            tt = E.rows_sumsq(ts.view(1, -1), n)
            v = E.gram(Yt, ts.view(1, -1), n)[:, 0].contiguous()
            scale_rows_(v.view(1, -1), q, tt, True)
            u = E.skinny_gemm(Yt, n, v.view(1, -1), [0, q], dense=True)[0].contiguous()
            normalize_(u, n)
            w = xt_vec(Xt, n, u, boff_dev, B).contiguous()  # :503-504
            normalize_over_features_(w, p, group)
            raw = blockprod(w)
        a = block_sumsq(w, boff_dev, B, group)  # :405-408
        wb = scale_by_block(w, boff_dev, B, a, p)
        tb = raw.clone()
        scale_rows_(tb, n, torch.sqrt(a), True)  # t_b = X_b w_b (:410-413)
        pv = xt_vec(Xt, n, ts, boff_dev, B, uu=tt)  # p = X'ts / ts'ts (:427)
        if rg is not None:
            pv = pv.contiguous()
            E.allreduce_(pv, rg)
        pss_dev[k] = block_sumsq(pv, boff_dev, B, group)
        tt_dev[k:k + 1] = tt
        vv_dev[k:k + 1] = E.rows_sumsq(v.view(1, -1), q)
        rank1_update_(Xt, n, ts, pv)  # X <- X - ts p' (:443)
        A_dev[k] = a
        Wc[k, :p], Wb[k, :p], P[k, :p], Ts[k], U[k], V[k] = w, wb, pv, ts, u, v
        Tb[:, k, :] = tb
    A = A_dev.cpu().numpy().T.copy()
    pssb_h, tt_h, vv_h = pss_dev.cpu().numpy(), tt_dev.cpu().numpy(), vv_dev.cpu().numpy()
    for k in range(K):  # explained variances (:429-448): ((ts p')**2).sum() == ts'ts * p'p
        evx.append(float(tt_h[k] * pssb_h[k].sum() / varxb.sum()))
        evxb[:, k] = tt_h[k] * pssb_h[k] / varxb
        evy.append(float(tt_h[k] * vv_h[k] / vary))
    Wc, Wb, P = Wc[:, :p], Wb[:, :p], P[:, :p]
    M = E.small_pinv(E.gram(P, Wc, p, group))
    R = E.right_multiply(Wc, p, None, M)  # :476
    beta = E.right_multiply(R, p, None, V.contiguous())  # :477
    model.A_ = A
    model.A_corrected_ = np.stack([_bip_corrected(A[:, k], shard.sizes) for k in range(K)], axis=1)
    model.explained_var_x_, model.explained_var_y_, model.explained_var_xblocks_ = evx, evy, evxb
    model.W_non_normal_ = [np.empty((s, 0)) for s in shard.sizes]
    model.W_concat_ = np.empty((pg, 0))
    _finish_common(model, shard, n, q, R, beta, P, Ts, U, V, Wb, Tb)
    model.__dict__["_cv_weights"] = Wc  # prefix models for model_selection.cross_val_predict


# ---- KERNEL (mbpls.py:576-807) -------------------------------------------------------------------
def crossprod(A, Bm, M, N, Kdim, kmajor, ldc):
    """FP64 tensor-core cross product; A is Bm (X'X / XX') -> SYRK: upper tiles only, then mirrored."""
    dev = A.device
    sym = 1 if (A.data_ptr() == Bm.data_ptr() and M == N) else 0
    if sym and os.environ.get("MBPLS_XP_SPLITS", "balanced") != "rectangular":
        splits = call("mbpls_crossprod_splits_syrk", M, Kdim)  # whole rounds of the CTAs on / above the diagonal
    else:
        splits = call("mbpls_crossprod_splits", M, N, Kdim)
    part = torch.zeros((splits, M * ldc), dtype=F64, device=dev)
    call("mbpls_crossprod_f64", ptr(A), A.stride(0), ptr(Bm), Bm.stride(0), M, N, Kdim, 1 if kmajor else 0, splits, ptr(part),
         ldc, sym, stream_ptr(dev))
    out = torch.zeros((M, ldc), dtype=F64, device=dev)
    call("mbpls_reduce_chunks_f64", ptr(part), splits, M * ldc, ptr(out), stream_ptr(dev))
    if sym:
        call("mbpls_symmetrize_f64", ptr(out), ldc, M, stream_ptr(dev))
    return out


def _fit_kernel(model, Xt, Yt, n, q, shard, boff_dev, zss, group, device, rows_group=None, n_global=None):
    """rows_group: the SAMPLE axis is sharded over this group (Xt / Yt hold this rank's n rows of all features);
    sums over samples are all-reduced (p x p + p x q partial cross-products), everything indexed by
    features is replicated, everything indexed by samples stays row-local."""
    mw = weakref.proxy(model)
    K, B = int(model.n_components), len(shard.sizes)
    p, ld = Xt.shape
    pg = shard.p_global
    rg = rows_group
    ng = n if n_global is None else n_global
    if group is not None and rg is None and ng >= pg:
        raise NotImplementedError("feature-sharded KERNEL covers p > n; n >= p uses the row-sharded path")

    def normalize_rows_(t):  # unit norm over the *global* sample axis
        if rg is None:
            return normalize_(t, n)
        nrm = torch.sqrt(E.rows_sumsq(t.view(1, -1), n, rg))
        scale_rows_(t.view(1, -1), n, nrm, True)
        return nrm
    st = stream_ptr(device)
    calc_all = bool(model.calc_all)
    V = torch.zeros((K, q), dtype=F64, device=device)
    P = torch.zeros((K, p), dtype=F64, device=device)
    Wc = torch.zeros((K, p), dtype=F64, device=device)
    Wb = torch.zeros((K, p), dtype=F64, device=device)
    U = torch.zeros((K, ld), dtype=F64, device=device)
    A_dev = torch.zeros((K, B), dtype=F64, device=device)
    if ng >= pg:  # Lindgren kernel (:580-650)
        COVt = xt_multi(Xt, n, Yt).contiguous()  # COVAR' : q x p (:587)
        ldv = (p + 15) // 16 * 16
        VAR = crossprod(Xt, Xt, p, p, n, True, ldv)  # X'X on the FP64 tensor cores (:586)
        E.allreduce_(COVt, rg)  # row-sharded: partial cross-products of this rank's samples
        E.allreduce_(VAR, rg)
        scal = torch.zeros(4, dtype=F64, device=device)
        for k in range(K):
            c = E.small_top_eigvec(E.gram(COVt, COVt, p))  # S = COVAR COVAR' (:584, :631)
            w = E.right_multiply(COVt, p, None, c.view(q, 1))[0].contiguous()
            normalize_over_features_(w, p, None)
            Vw = torch.zeros(p, dtype=F64, device=device)
            call("mbpls_dense_gemv_f64", ptr(VAR), ldv, p, p, ptr(w), ptr(Vw), st)
            den = E.gram(w.view(1, -1), Vw.view(1, -1), p)  # w'VAR w (:595)
            scal[0:1] = den.view(-1)
            wC = E.gram(COVt, w.view(1, -1), p)[:, 0].contiguous()  # w'COVAR  (= den * v)
            v, pv = wC.clone(), Vw.clone()
            scale_rows_(v.view(1, -1), q, scal, True)  # v = w'COVAR / den (:596)
            scale_rows_(pv.view(1, -1), p, scal, True)  # p = w'VAR / den (:599)
            if calc_all:  # :601-627
                a = block_sumsq(w, boff_dev, B, None)
                Wb[k] = scale_by_block(w, boff_dev, B, a, p)
                A_dev[k] = a
                u = E.skinny_gemm(Yt, n, v.view(1, -1), [0, q], dense=True)[0].contiguous()  # Y v' / (v v') up to the scale ...
                normalize_rows_(u)  # ... which the normalisation removes (:624-625)
                U[k] = u
            # COVAR <- D'COVAR, VAR <- D'VAR D = VAR - den p'p  (:630-633)
            call("mbpls_rank1_update_f64", ptr(COVt), COVt.stride(0), p, q, ptr(pv), ptr(wC), st)
            call("mbpls_dense_rank2_f64", ptr(VAR), ldv, p, p, ptr(pv), ptr(pv), ptr(scal), 0.0, 0.0, -1.0, -1, 0, st)
            V[k], P[k], Wc[k] = v, pv, w
        M = E.small_pinv(E.gram(P, Wc, p))
        R = E.right_multiply(Wc, p, None, M)  # :642
        beta = E.right_multiply(R, p, None, V.contiguous())  # :643
        Ts = E.skinny_gemm(Xt, n, R, shard.block_off, dense=True)  # Ts = X R (:644)
        nrm = torch.sqrt(E.rows_sumsq(Ts, n, rg))  # :646-650
        scale_rows_(V, q, nrm, False)
        scale_rows_(P, p, nrm, False)
        scale_rows_(Ts, n, nrm, True)
    else:  # Rannar kernel (:694-738)
        AX = crossprod(Xt, Xt, n, n, p, False, ld)  # X X' on the FP64 tensor cores (:704); features sharded:
        E.allreduce_(AX, group)                     # one n x n all-reduce of the partial association matrices
        Yc = Yt.clone()
        Ts = torch.zeros((K, ld), dtype=F64, device=device)
        scal = torch.zeros(4, dtype=F64, device=device)
        for k in range(K):
            At = E.skinny_gemm(AX[:, :], n, Yc, [0, n], dense=True)  # (AS_X Y)' : q x ld;  S = AS_X AS_Y = (AS_X Y) Y' (:707, :725)
            c = E.small_top_sv_product(E.gram(Yc, Yc, n), E.gram(At, At, n))
            ts = E.right_multiply(At, ld, None, c.view(q, 1))[0].contiguous()
            ts[n:] = 0.0
            normalize_(ts, n)  # :713
            yt = E.gram(Yc, ts.view(1, -1), n)[:, 0].contiguous()  # Y'ts
            u = E.skinny_gemm(Yc, n, yt.view(1, -1), [0, q], dense=True)[0].contiguous()  # AS_Y ts (:718)
            normalize_(u, n)
            kv = torch.zeros(ld, dtype=F64, device=device)
            call("mbpls_dense_gemv_f64", ptr(AX), ld, n, n, ptr(ts), ptr(kv), st)
            scal[0:1] = E.gram(ts.view(1, -1), kv.view(1, -1), n).view(-1)  # ts' AS_X ts
            # AS_X <- D AS_X D,  Y <- D Y  with D = I - ts ts' (:722-724)
            call("mbpls_dense_rank2_f64", ptr(AX), ld, n, n, ptr(ts), ptr(kv), ptr(scal), -1.0, -1.0, 1.0, -1, 0, st)
            call("mbpls_rank1_update_f64", ptr(Yc), ld, n, q, ptr(ts), ptr(yt), st)
            U[k], Ts[k] = u, ts
        Wc = xt_multi(Xt, n, U).contiguous()  # X'U (:731)
        wn = torch.sqrt(E.rows_sumsq(Wc, p, group))
        if p > 0:
            scale_rows_(Wc, p, wn, True)  # :733
        Gt = E.small_pinv(E.gram(Ts, Ts, n))  # (Ts'Ts)^+, K x K
        XtTs = xt_multi(Xt, n, Ts).contiguous()
        P = E.right_multiply(XtTs, p, None, Gt).contiguous()  # :734
        V = E.right_multiply(E.gram(Ts, Yt, n).contiguous(), q, None, Gt).contiguous()  # Y'Ts (Ts'Ts)^+ as K x q (:735)
        M = E.small_pinv(E.gram(P, Wc, p, group))
        R = E.right_multiply(Wc, p, None, M)  # :737
        beta = E.right_multiply(R, p, None, V.contiguous())  # :738
    Tb = None
    evx, evy, evxb = [], [], np.empty((B, 0))
    if calc_all:  # :653-689 / :740-802
        if zss is None:
            zss = E.feature_sumsq(Xt, n)
        varxb = model._block_sums(zss, boff_dev, B, group if rg is None else rg)
        vary_t = E.segsum(E.feature_sumsq(Yt, n), E._i32([0, q], device), 1).clone()
        E.allreduce_(vary_t, rg)
        vary = float(vary_t.item())
        blockprod = BlockProducts(Xt, n, shard.block_off, group)
        Tb = torch.zeros((B, K, ld), dtype=F64, device=device)
        evxb = np.zeros((B, K))
        pss_dev = torch.zeros((K, B), dtype=F64, device=device)
        for k in range(K):
            if n < pg:
                a = block_sumsq(Wc[k].contiguous(), boff_dev, B, group)
                Wb[k] = scale_by_block(Wc[k].contiguous(), boff_dev, B, a, p)
                A_dev[k] = a
            Tb[:, k, :] = blockprod(Wb[k].contiguous())
            pss_dev[k] = block_sumsq(P[k].contiguous(), boff_dev, B, group)
            rank1_update_(Xt, n, Ts[k].contiguous(), P[k].contiguous())
        tt_h = E.rows_sumsq(Ts, n, rg).cpu().numpy()
        vv_h = E.rows_sumsq(V.contiguous(), q).cpu().numpy()
        pssb_h = pss_dev.cpu().numpy()
        for k in range(K):  # ((Ts_k P_k')**2).sum() == Ts_k'Ts_k * P_k'P_k (:669-689)
            evx.append(float(tt_h[k] * pssb_h[k].sum() / varxb.sum()))
            evxb[:, k] = tt_h[k] * pssb_h[k] / varxb
            evy.append(float(tt_h[k] * vv_h[k] / vary))
        A = A_dev.cpu().numpy().T.copy()
        model.A_ = A
        model.A_corrected_ = np.stack([_bip_corrected(A[:, k], shard.sizes) for k in range(K)], axis=1)
    else:
        model.A_ = np.empty((B, 0))
        model.A_corrected_ = np.empty((B, 0))
    model.explained_var_x_, model.explained_var_y_, model.explained_var_xblocks_ = evx, evy, evxb
    model.W_non_normal_ = [np.empty((s, 0)) for s in shard.sizes]
    Wc_keep = Wc
    extra = {"W_concat_": lambda: mw._features_T(Wc_keep, shard, False)}
    if not calc_all:
        extra["W_"] = lambda: [np.empty((s, 0)) for s in shard.sizes]
        extra["T_"] = lambda: [np.empty((ng, 0)) for _ in shard.sizes]
        extra["U_"] = (lambda: np.empty((ng, 0))) if ng >= pg else (lambda: mw._gather_samples(U, n))
    _finish_common(model, shard, n, q, R, beta, P, Ts, U, V, Wb if calc_all else None, Tb, extra)
    model.__dict__["_cv_weights"] = Wc  # prefix models: R_k = W_k pinv(P_k'W_k) (the column scaling of P_ / V_ cancels in beta)
