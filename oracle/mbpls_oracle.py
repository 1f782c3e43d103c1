"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference MB-PLS fit path.

All ``file:line`` citations are relative to ``/root/reference`` and point at the
region of ``mbpls/mbpls.py`` (v1.0.4) each function follows.  The code is a
restatement, not a copy: the reference is one 1,500-line method with four
inlined algorithm bodies and Python loops over NaN rows/columns; here every
step is a small vectorised function so that (a) it runs fast enough to be the
CPU baseline and (b) the CUDA kernels can be compared step by step.

Third-party arithmetic that the reference calls but that is *not* in
``/root/reference``:

* scikit-learn ``StandardScaler`` (``mbpls/mbpls.py:307,314,325-326``; minimum
  version 0.22.1 per ``setup.py:25``, nothing pinned).  ``OracleScaler``
  restates its published algorithm (nan-aware corrected two-pass variance,
  population std, near-constant features get scale 1); tests compare it with the
  installed scikit-learn.
* numpy / LAPACK ``svd`` / ``pinv`` and scipy ``svds`` (ARPACK, random start
  vector) for the top singular vector.  The oracle always uses the dense
  ``np.linalg.svd`` -- ``svds`` only differs by sign and ARPACK tolerance, and
  the reference's own tests compare absolute values.

Parity pin: see ``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import warnings

import numpy as np

_EPS = np.finfo(np.float64).eps


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def legacy_ortho_group_rvs(dim: int, random_state: int) -> np.ndarray:
    """Householder-loop ``scipy.stats.ortho_group.rvs`` as shipped in 2018.

    The reference's known-answer CSVs were generated with this sampler
    (``mbpls/tests/test_mbpls.py:41-42``); today's scipy draws a different
    matrix from the same seed.
    """
    rs = np.random.RandomState(random_state)
    H = np.eye(dim)
    for n in range(dim):
        x = rs.normal(size=(dim - n,))
        norm2 = x @ x
        x0 = x[0].item()
        D = np.sign(x[0]) if x[0] != 0 else 1
        x[0] += D * np.sqrt(norm2)
        x /= np.sqrt((norm2 - x0 ** 2 + x[0] ** 2) / 2.0)
        H[:, n:] = -D * (H[:, n:] - np.outer(H[:, n:] @ x, x))
    return H


def nan_census(data: np.ndarray):
    """NaN census of one 2-D array -- follows ``mbpls/mbpls.py:255-271``.

    Returns ``(sparse_rows, sparse_columns, dense_rows, dense_columns)`` as
    int64 index arrays and emits the reference's >50 % warnings.
    """
    nan = np.isnan(data)
    col_has = nan.any(axis=0)
    row_has = nan.any(axis=1)
    msg = ("The sparsity of your data is likely to high for this algorithm. This can cause either convergence"
           "problems or crash the algorithm.")
    if col_has.sum() / data.shape[1] > 0.5:
        warnings.warn(msg)
    if row_has.sum() / data.shape[0] > 0.5:
        warnings.warn(msg)
    return (np.where(row_has)[0], np.where(col_has)[0],
            np.where(~row_has)[0], np.where(~col_has)[0])


class OracleScaler:
    """Restatement of scikit-learn ``StandardScaler(with_mean=True, with_std=True)``.

    Used by the reference at ``mbpls/mbpls.py:307,314`` (X blocks) and
    ``:325-326`` (Y).  Algorithm (scikit-learn ``_incremental_mean_and_var`` on a
    first batch + ``_is_constant_feature``): nan-aware column sum / observed
    count -> mean; corrected two-pass variance (Chan, Golub & LeVeque) with
    ddof=0; scale = sqrt(var) except for columns whose variance is within the
    two-pass error bound of zero, which get scale 1.  NaNs pass through.
    """

    def fit(self, X: np.ndarray) -> "OracleScaler":
        X = np.asarray(X, dtype=np.float64)
        nan = np.isnan(X)
        has_nan = bool(nan.any())
        sum_op = np.nansum if has_nan else np.sum
        n = X.shape[0]
        col_sum = sum_op(X, axis=0)
        count = n - sum_op(nan.astype(np.float64), axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = col_sum / count
            centred = X - mean
            corr = sum_op(centred, axis=0)
            centred **= 2
            ssq = sum_op(centred, axis=0)
            ssq -= corr ** 2 / count
            var = ssq / count
        self.mean_ = mean
        self.var_ = var
        self.n_samples_seen_ = count.astype(np.int64)
        if np.ptp(self.n_samples_seen_) == 0:  # sklearn collapses a constant count to a scalar
            self.n_samples_seen_ = self.n_samples_seen_[0]
        bound = count * _EPS * var + (count * mean * _EPS) ** 2
        scale = np.sqrt(var)
        scale[var <= bound] = 1.0
        self.scale_ = scale
        return self

    def transform(self, X: np.ndarray) -> np.ndarray:
        Z = np.array(X, dtype=np.float64, copy=True)
        Z -= self.mean_
        Z /= self.scale_
        return Z

    def fit_transform(self, X: np.ndarray) -> np.ndarray:
        return self.fit(X).transform(X)

    def inverse_transform(self, Z: np.ndarray) -> np.ndarray:
        X = np.array(Z, dtype=np.float64, copy=True)
        X *= self.scale_
        X += self.mean_
        return X


def _unit(x: np.ndarray) -> np.ndarray:
    return x / np.linalg.norm(x)


def _top_left_singular_vector(S: np.ndarray) -> np.ndarray:
    """``np.linalg.svd(S)[0][:, 0:1]`` -- ``mbpls/mbpls.py:398,491,590,713,1001``."""
    return np.linalg.svd(S, full_matrices=False)[0][:, 0:1]


def _zero_nan(A: np.ndarray) -> np.ndarray:
    return np.where(np.isnan(A), 0.0, A)


def _bip_corrected(a, sizes):
    """Block importance corrected for block size -- ``mbpls/mbpls.py:450-461`` (same text
    at ``:542-553``, ``:611-622``, ``:790-801``, ``:950-963``)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    if a.size == 1:
        return np.array([1.0])
    sizes = np.asarray(sizes, dtype=np.float64)
    corrected = a * (1.0 - sizes / sizes.sum())
    return corrected / corrected.sum()


def _masked_regress_cols(A, s, sparse_cols, plain_den):
    """Column-wise regression of A (n x m) on s (n x 1).

    Fully observed columns: ``A[:, j].s / plain_den``; columns listed in
    ``sparse_cols``: sum over observed rows of ``a_ij s_i`` divided by the sum of
    ``s_i**2`` over the same rows.  ``mbpls/mbpls.py:846-854`` (X weights),
    ``:890-897`` (Y weights), ``:919-926`` (loadings, where plain_den is None
    because the reference does not divide dense columns, ``:920``).
    """
    A0 = _zero_nan(A)
    num = A0.T @ s
    out = num if plain_den is None else num / plain_den
    if len(sparse_cols):
        obs = (~np.isnan(A[:, sparse_cols])).astype(np.float64)
        den = obs.T @ (s ** 2)
        out = out.copy()
        out[sparse_cols] = num[sparse_cols] / den
    return out


def _masked_regress_rows(A, w, sparse_rows, plain_den):
    """Row-wise regression of A (n x m) on w (m x 1) -- ``mbpls/mbpls.py:865-873`` and
    ``:901-907``.  Rows in ``sparse_rows`` use observed columns only."""
    A0 = _zero_nan(A)
    full = A @ w  # NaN rows stay NaN exactly as in the reference (:866, :902)
    out = full if plain_den is None else full / plain_den
    if len(sparse_rows):
        sub = A[sparse_rows]
        obs = (~np.isnan(sub)).astype(np.float64)
        num = A0[sparse_rows] @ w
        den = obs @ (w ** 2)
        out = out.copy()
        out[sparse_rows] = num / den
    return out


# --------------------------------------------------------------------------- #
# the estimator
# --------------------------------------------------------------------------- #
class OracleMBPLS:
    """CPU oracle with the constructor and attribute contract of ``mbpls.mbpls.MBPLS``
    (``mbpls/mbpls.py:243-253``).  Extra, oracle-only attribute: ``n_iter_`` (NIPALS
    trips per component; the reference keeps it in a local, ``:839,914``)."""

    def __init__(self, n_components=2, full_svd=False, method='NIPALS', standardize=True, max_tol=1e-14,
                 nipals_convergence_norm=2, calc_all=True, sparse_data=False, copy=True, max_iter=100000):
        self.n_components = n_components
        self.full_svd = full_svd
        self.method = method
        self.standardize = standardize
        self.max_tol = max_tol
        self.nipals_convergence_norm = nipals_convergence_norm
        self.calc_all = calc_all
        self.sparse_data = sparse_data
        self.copy = copy
        self.max_iter = max_iter  # safety cap, not in the reference (its loop is unbounded, :841)

    # ---- prologue: mbpls.py:286-382 --------------------------------------
    def _ingest(self, X, Y, fitting: bool):
        if isinstance(X, list) and not isinstance(X[0], list):
            blocks = [np.array(b, dtype=np.float64, copy=True) for b in X]
        else:
            blocks = [np.array(X, dtype=np.float64, copy=True)]
        for b in blocks:
            if b.ndim != 2:
                raise ValueError("each X block must be 2-D")
            if not self.sparse_data and not np.isfinite(b).all():
                raise ValueError("Input X contains NaN or infinity.")
        if Y is not None:
            Y = np.array(Y, dtype=np.float64, copy=True)
            if not self.sparse_data and not np.isfinite(Y).all():
                raise ValueError("Input Y contains NaN or infinity.")
        return blocks, Y

    def fit(self, X, Y):
        if self.sparse_data is True and self.method != 'NIPALS':  # :286-290
            warnings.warn("The parameter sparse data was set to 'True', but the chosen method is not 'NIPALS'."
                          "The method will be set to 'NIPALS'")
            self.method = 'NIPALS'
        if self.method not in ('NIPALS', 'UNIPALS', 'KERNEL', 'SIMPLS'):
            raise NameError('Method you called is unknown')  # :1050
        blocks, Y = self._ingest(X, Y, True)
        if Y.ndim == 1:
            Y = Y.reshape(-1, 1)
        if self.sparse_data:
            # :294-296.  The reference takes the census before the 1-D reshape (:297-298) and therefore
            # only accepts 2-D Y in NaN mode; for 2-D Y the result is identical.
            self.sparse_Y_info_ = {'Y': nan_census(Y)}
        for b in blocks:
            if b.shape[0] != Y.shape[0]:
                raise ValueError("Found input variables with inconsistent numbers of samples")
        if self.sparse_data:
            self.sparse_X_info_ = {i: nan_census(b) for i, b in enumerate(blocks)}  # :313,321
        if self.standardize:  # :299-326
            self.x_scalers_ = [OracleScaler() for _ in blocks]
            blocks = [s.fit_transform(b) for s, b in zip(self.x_scalers_, blocks)]
            self.y_scaler_ = OracleScaler()
            Y = self.y_scaler_.fit_transform(Y)
        self.num_blocks_ = len(blocks)
        sizes = [b.shape[1] for b in blocks]
        bounds = np.concatenate(([0], np.cumsum(sizes)))
        self._bounds = bounds
        n, q, B, p = Y.shape[0], Y.shape[1], len(blocks), int(bounds[-1])
        # :359-382 empty containers
        self.W_ = [np.empty((s, 0)) for s in sizes]
        self.W_non_normal_ = [np.empty((s, 0)) for s in sizes]
        if self.method != 'SIMPLS':
            self.A_ = np.empty((B, 0))
            self.A_corrected_ = np.empty((B, 0))
            self.T_ = [np.empty((n, 0)) for _ in sizes]
        self.explained_var_xblocks_ = np.empty((B, 0))
        self.V_ = np.empty((q, 0))
        self.U_ = np.empty((n, 0))
        self.Ts_ = np.empty((n, 0))
        self.explained_var_y_ = []
        self.explained_var_x_ = []
        self.P_ = np.empty((p, 0))
        self.W_concat_ = np.empty((p, 0))
        Xc = np.hstack(blocks)  # :379
        getattr(self, '_fit_' + self.method.lower())(Xc, Y, sizes)
        return self

    def _split(self, M):
        return [M[self._bounds[b]:self._bounds[b + 1]] for b in range(self.num_blocks_)]

    # ---- NIPALS: mbpls.py:809-993 ----------------------------------------
    def _fit_nipals(self, X, Y, sizes):
        B, bounds, sparse = self.num_blocks_, self._bounds, bool(self.sparse_data)
        Xb = [X[:, bounds[b]:bounds[b + 1]] for b in range(B)]  # :816-818
        n = X.shape[0]
        if sparse:
            xinfo = self.sparse_X_info_
            yinfo = self.sparse_Y_info_['Y']
        self.n_iter_ = []
        self.diff_trace_ = []
        P_cols, Wn_cols, Wt_cols, T_cols = [], [], [], []
        for comp in range(self.n_components):
            if self.calc_all and comp == 0:  # :822-830
                varx = np.nansum(X ** 2) if sparse else (X ** 2).sum()
                vary = np.nansum(Y ** 2) if sparse else (Y ** 2).sum()
            if sparse:  # :832-836
                if len(yinfo[1]) == Y.shape[1]:
                    u = np.random.rand(n, 1)
                else:
                    c0 = yinfo[3][0]
                    u = Y[:, c0:c0 + 1]
            else:
                u = Y[:, 0:1]  # :838
            run, diff, trace = 1, 1.0, []
            ts_old = None
            while diff > self.max_tol and run <= self.max_iter:  # :841
                uu = u.T @ u
                wt, w = [], []
                for b in range(B):  # :845-860
                    if sparse:
                        wb = _masked_regress_cols(Xb[b], u, xinfo[b][1], uu)
                    else:
                        wb = Xb[b].T @ u / uu
                    wt.append(wb)
                    w.append(wb / np.linalg.norm(wb))
                t = []
                for b in range(B):  # :862-875
                    if sparse:
                        t.append(_masked_regress_rows(Xb[b], w[b], xinfo[b][0], None))
                    else:
                        t.append(Xb[b] @ w[b])
                T = np.hstack(t)  # :877
                a = _unit(T.T @ u / uu)  # :879-880
                ts = _unit(T @ a)  # :882-883
                if run > 1:  # :884-887  (matrix norm of an n x 1 array)
                    diff = np.linalg.norm(ts_old - ts, ord=self.nipals_convergence_norm)
                    trace.append(float(diff))
                ts_old = ts.copy()
                tt = ts.T @ ts
                if sparse:  # :890-897
                    v = _masked_regress_cols(Y, ts, yinfo[1], tt)
                else:
                    v = Y.T @ ts / tt  # :899
                vv = v.T @ v
                if sparse:
                    # :901-909 -- NB the reference loops over the sparse rows of the *last X block*
                    # (`self.sparse_X_info_[block][0]` with `block` left over from the loop at :863).
                    # Those rows divide by the masked v'v (:906-907), all others by the plain v'v (:902).
                    uraw = _masked_regress_rows(Y, v, xinfo[B - 1][0], vv)
                else:
                    uraw = Y @ v / vv  # :911
                u = _unit(uraw)
                run += 1
            self.n_iter_.append(run - 1)
            self.diff_trace_.append(trace)
            # loadings :917-930
            if sparse:
                pl = [_masked_regress_cols(Xb[b], ts, xinfo[b][1], None) for b in range(B)]
                # masked columns divide by the masked ts'ts (:923-925); dense columns are plain X'ts (:920)
            else:
                pl = [Xb[b].T @ ts for b in range(B)]
            p_tot = np.vstack(pl)
            a2 = a ** 2
            if self.calc_all:  # :932-964 ; ((ts p')**2).sum() == (ts'ts)(p'p)
                tt = float((ts ** 2).sum())
                self.explained_var_x_.append(tt * float((p_tot ** 2).sum()) / varx)
                self.explained_var_y_.append(tt * float((v ** 2).sum()) / vary)
                if comp == 0:
                    varxb = [np.nansum(Xb[b] ** 2) if sparse else (Xb[b] ** 2).sum() for b in range(B)]
                col = [tt * float((pl[b] ** 2).sum()) / varxb[b] for b in range(B)]
                self.explained_var_xblocks_ = np.hstack((self.explained_var_xblocks_, np.array(col).reshape(-1, 1)))
                self.A_corrected_ = np.hstack((self.A_corrected_, _bip_corrected(a2, sizes).reshape(-1, 1)))
            Xb = [Xb[b] - ts @ pl[b].T for b in range(B)]  # :968-969 (Y is not deflated, :971-972)
            self.V_ = np.hstack((self.V_, v))  # :975-983
            self.U_ = np.hstack((self.U_, u))
            self.A_ = np.hstack((self.A_, a2))
            self.Ts_ = np.hstack((self.Ts_, ts))
            P_cols.append(p_tot)
            Wn_cols.append(w)
            Wt_cols.append(wt)
            T_cols.append(t)
        self.P_ = np.hstack(P_cols)
        for b in range(B):
            self.W_[b] = np.hstack([w[b] for w in Wn_cols])
            self.W_non_normal_[b] = np.hstack([w[b] for w in Wt_cols])
            self.T_[b] = np.hstack([t[b] for t in T_cols])
        Wtot = np.concatenate(self.W_non_normal_, axis=0)  # :986-989
        Wtot = Wtot / np.linalg.norm(Wtot, axis=0)
        self.R_ = Wtot @ np.linalg.pinv(self.P_.T @ Wtot)
        self.beta_ = self.R_ @ self.V_.T
        self.P_ = self._split(self.P_)  # :991

    # ---- shared tail of UNIPALS: mbpls.py:402-477 / :495-570 ---------------
    def _block_parts(self, w_full):
        w, a = [], []
        for part in self._split(w_full):  # :405-408
            nrm = np.linalg.norm(part)
            w.append(part / nrm)
            a.append(nrm ** 2)
        return w, a

    def _fit_unipals(self, X, Y, sizes):
        n, p = X.shape
        B, bounds = self.num_blocks_, self._bounds
        W_cols = []
        vary = (Y ** 2).sum()
        for comp in range(self.n_components):
            if n >= p:  # :388-424
                S = np.linalg.multi_dot([X.T, Y, Y.T, X])
                wfull = _top_left_singular_vector(S)
                ts = _unit(X @ wfull)
                v = Y.T @ ts / (ts.T @ ts)
                u = _unit(Y @ v)
            else:  # :481-504
                S = np.linalg.multi_dot([X, X.T, Y, Y.T])
                ts = _top_left_singular_vector(S)
                v = Y.T @ ts / (ts.T @ ts)
                u = _unit(Y @ v)
                wfull = _unit(X.T @ u)
            w, a = self._block_parts(wfull)
            t = [X[:, bounds[b]:bounds[b + 1]] @ w[b] for b in range(B)]  # :410-413 / :514-517
            tt = float(ts.T @ ts)
            pl = X.T @ ts / tt  # :427 / :520
            if comp == 0:
                varx = (X ** 2).sum()
                varxb = [(X[:, bounds[b]:bounds[b + 1]] ** 2).sum() for b in range(B)]
            self.explained_var_x_.append(tt * float((pl ** 2).sum()) / varx)
            col = [tt * float((pl[bounds[b]:bounds[b + 1]] ** 2).sum()) / varxb[b] for b in range(B)]
            X = X - ts @ pl.T  # :443 / :535
            self.explained_var_y_.append(tt * float((v ** 2).sum()) / vary)
            self.V_ = np.hstack((self.V_, v))
            self.U_ = np.hstack((self.U_, u))
            self.A_ = np.hstack((self.A_, np.array(a).reshape(-1, 1)))
            self.A_corrected_ = np.hstack((self.A_corrected_, _bip_corrected(a, sizes).reshape(-1, 1)))
            self.explained_var_xblocks_ = np.hstack((self.explained_var_xblocks_, np.array(col).reshape(-1, 1)))
            self.Ts_ = np.hstack((self.Ts_, ts))
            self.P_ = np.hstack((self.P_, pl))
            W_cols.append(wfull)
            for b in range(B):
                self.W_[b] = np.hstack((self.W_[b], w[b]))
                self.T_[b] = np.hstack((self.T_[b], t[b]))
        weights = np.hstack(W_cols)
        self.R_ = weights @ np.linalg.pinv(self.P_.T @ weights)  # :476-477 (last component's value survives)
        self.beta_ = self.R_ @ self.V_.T
        self.P_ = self._split(self.P_)

    # ---- KERNEL: mbpls.py:576-807 ------------------------------------------
    def _fit_kernel(self, X, Y, sizes):
        n, p = X.shape
        B, bounds = self.num_blocks_, self._bounds
        if n >= p:  # :580-650
            S = np.linalg.multi_dot([X.T, Y, Y.T, X])
            VAR = X.T @ X
            COVAR = X.T @ Y
            for comp in range(self.n_components):
                w = _top_left_singular_vector(S)
                den = np.linalg.multi_dot([w.T, VAR, w])
                v = w.T @ COVAR / den
                pl = w.T @ VAR / den
                if self.calc_all:  # :601-627
                    wb, a = self._block_parts(w)
                    for b in range(B):
                        self.W_[b] = np.hstack((self.W_[b], wb[b]))
                    self.A_corrected_ = np.hstack((self.A_corrected_, _bip_corrected(a, sizes).reshape(-1, 1)))
                    self.U_ = np.hstack((self.U_, _unit(Y @ v.T / (v @ v.T))))
                    self.A_ = np.hstack((self.A_, np.array(a).reshape(-1, 1)))
                D = np.eye(p) - w @ pl  # :630-633
                S = np.linalg.multi_dot([D.T, S, D])
                VAR = np.linalg.multi_dot([D.T, VAR, D])
                COVAR = D.T @ COVAR
                self.V_ = np.hstack((self.V_, v.T))
                self.P_ = np.hstack((self.P_, pl.T))
                self.W_concat_ = np.hstack((self.W_concat_, w))
            self.R_ = self.W_concat_ @ np.linalg.pinv(self.P_.T @ self.W_concat_)  # :642-650
            self.beta_ = self.R_ @ self.V_.T
            self.Ts_ = X @ self.R_
            nrm = np.linalg.norm(self.Ts_, axis=0)
            self.V_ = self.V_ * nrm
            self.P_ = self.P_ * nrm
            self.Ts_ = self.Ts_ / nrm
        else:  # :694-738
            AX = X @ X.T
            AY = Y @ Y.T
            S = AX @ AY
            for comp in range(self.n_components):
                ts = _top_left_singular_vector(S)
                u = _unit(AY @ ts)
                D = np.eye(n) - ts @ ts.T
                AX = np.linalg.multi_dot([D, AX, D])
                AY = np.linalg.multi_dot([D, AY, D])
                S = AX @ AY
                self.U_ = np.hstack((self.U_, u))
                self.Ts_ = np.hstack((self.Ts_, ts))
            Wc = X.T @ self.U_
            self.W_concat_ = Wc / np.linalg.norm(Wc, axis=0)
            G = np.linalg.pinv(self.Ts_.T @ self.Ts_)
            self.P_ = np.linalg.multi_dot([X.T, self.Ts_, G])
            self.V_ = np.linalg.multi_dot([Y.T, self.Ts_, G])
            self.R_ = self.W_concat_ @ np.linalg.pinv(self.P_.T @ self.W_concat_)
            self.beta_ = self.R_ @ self.V_.T
        if self.calc_all:  # :653-689 / :740-802
            varx = (X ** 2).sum()
            vary = (Y ** 2).sum()
            varxb = [(X[:, bounds[b]:bounds[b + 1]] ** 2).sum() for b in range(B)]
            for k in range(self.n_components):
                if n < p:  # block weights come from W_concat_ (:746-752)
                    wb, a = self._block_parts(self.W_concat_[:, k:k + 1])
                    self.A_ = np.hstack((self.A_, np.array(a).reshape(-1, 1)))
                    self.A_corrected_ = np.hstack((self.A_corrected_, _bip_corrected(a, sizes).reshape(-1, 1)))
                    for b in range(B):
                        self.W_[b] = np.hstack((self.W_[b], wb[b]))
                for b in range(B):
                    tb = X[:, bounds[b]:bounds[b + 1]] @ self.W_[b][:, k:k + 1]
                    self.T_[b] = np.hstack((self.T_[b], tb))
                tk, pk, vk = self.Ts_[:, k:k + 1], self.P_[:, k:k + 1], self.V_[:, k:k + 1]
                tt = float((tk ** 2).sum())
                self.explained_var_x_.append(tt * float((pk ** 2).sum()) / varx)
                col = [tt * float((pk[bounds[b]:bounds[b + 1]] ** 2).sum()) / varxb[b] for b in range(B)]
                self.explained_var_xblocks_ = np.hstack((self.explained_var_xblocks_, np.array(col).reshape(-1, 1)))
                X = X - tk @ pk.T
                self.explained_var_y_.append(tt * float((vk ** 2).sum()) / vary)
        self.P_ = self._split(self.P_)

    # ---- SIMPLS: mbpls.py:995-1048 -----------------------------------------
    def _fit_simpls(self, X, Y, sizes):
        warnings.warn("Method 'SIMPLS' does not calculate A_ and T_!")
        S = X.T @ Y  # :998
        R, T, P, Q, U, V, W = [], [], [], [], [], [], []
        for comp in range(self.n_components):
            qv = _top_left_singular_vector(S.T @ S)  # :1000-1003
            r = S @ qv
            w = r.copy()
            t = X @ r
            t = t - np.mean(t)  # :1007
            normt = np.sqrt(t.T @ t)
            t = t / normt
            r = r / normt
            pl = X.T @ t
            qv = Y.T @ t
            u = Y @ qv
            v = pl
            if comp > 0:  # :1015-1017
                Vm, Tm = np.hstack(V), np.hstack(T)
                v = v - Vm @ (Vm.T @ pl)
                u = u - Tm @ (Tm.T @ u)
            v = v / np.sqrt(v.T @ v)
            S = S - v @ (v.T @ S)  # :1019
            R.append(r); T.append(t); P.append(pl); Q.append(qv)
            U.append(u / np.linalg.norm(u)); V.append(v); W.append(w)
        self.P_ = self._split(np.hstack(P))
        self.Ts_ = np.hstack(T)
        self.U_ = np.hstack(U)
        self.R_ = np.hstack(R)
        self.V_ = np.hstack(Q)
        self.beta_ = self.R_ @ self.V_.T
        self.W_ = np.hstack(W)  # a single p x K array, not a list (:1046)

    # ---- transform: mbpls.py:1052-1335 --------------------------------------
    def transform(self, X, Y=None, return_block_scores=False, copy=True):
        if not hasattr(self, 'beta_'):
            raise AttributeError("not fitted")
        blocks, Y = self._ingest(X, Y, False)
        if self.standardize:
            blocks = [s.transform(b) for s, b in zip(self.x_scalers_, blocks)]
        Xc = np.hstack(blocks)
        Ts = _zero_nan(Xc) @ self.R_  # :1110-1117 (NaN rows: observed columns only == zero fill)
        out = [Ts]
        if self.method != 'SIMPLS' and return_block_scores:  # :1126-1155
            T = []
            for b in range(self.num_blocks_):
                Xb = blocks[b]
                cols = []
                for k in range(self.n_components):
                    if k > 0:
                        Xb = Xb - Ts[:, k - 1:k] @ self.P_[b][:, k - 1:k].T
                    cols.append(_zero_nan(Xb) @ self.W_[b][:, k:k + 1])
                T.append(np.hstack(cols))
            out.append(T)
        if Y is not None:  # :1119-1125, :1156-1166
            if Y.ndim == 1:
                Y = Y.reshape(-1, 1)
            if self.standardize:
                Y = self.y_scaler_.transform(Y)
            Uraw = _zero_nan(Y) @ self.V_
            out.append(Uraw / np.linalg.norm(Uraw, axis=0))
        return out[0] if len(out) == 1 else tuple(out)

    # ---- predict: mbpls.py:1337-1410 ----------------------------------------
    def predict(self, X, copy=True):
        if not hasattr(self, 'beta_'):
            raise AttributeError("not fitted")
        blocks, _ = self._ingest(X, None, False)
        if self.standardize:
            blocks = [s.transform(b) for s, b in zip(self.x_scalers_, blocks)]
        yhat = _zero_nan(np.hstack(blocks)) @ self.beta_  # :1379-1386
        if self.standardize:
            yhat = self.y_scaler_.inverse_transform(yhat)
        return yhat
