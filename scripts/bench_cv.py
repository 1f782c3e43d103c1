"""Leave-one-out model selection as the reference's notebooks run it (Carbohydrate_Microarray_PLS.ipynb: cross_val_predict(
MBPLS(n_components=k), X, y, cv=len(X)) for k = 1..15) on a 100 x (200 + 250) problem: all folds and all component counts in one
kernel launch, the per-fold loop on the streaming kernels, and -- for scale -- the numpy oracle's refit loop for ONE k."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sklearn.model_selection import LeaveOneOut
from mbpls_b200 import MBPLS
from mbpls_b200.model_selection import cross_val_predict
from oracle import OracleMBPLS
from oracle.cases import latent_blocks

warnings.simplefilter("ignore")
X, Y = latent_blocks(100, (200, 250), 1, 15, seed=7)
y = Y.ravel()
ks = list(range(1, 16))


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out

ta, a = t(lambda: cross_val_predict(MBPLS().set_runtime(small_path=True), X, y, cv=LeaveOneOut(), n_components_list=ks))
tb, b = t(lambda: cross_val_predict(MBPLS().set_runtime(small_path=False), X, y, cv=LeaveOneOut(), n_components_list=ks), reps=1)
t0 = time.perf_counter()
want = np.full(100, np.nan)
for tr, te in LeaveOneOut().split(np.arange(100)):
    m = OracleMBPLS(n_components=15).fit([x[tr] for x in X], y[tr])
    want[te] = m.predict([x[te] for x in X]).ravel()
tc = time.perf_counter() - t0
err = float(np.linalg.norm(a[15] - want) / np.linalg.norm(want))
print(json.dumps(dict(case="LOO 100 x (200+250), k = 1..15", one_launch_s=ta, fold_loop_streaming_s=tb,
                      oracle_numpy_one_k_s=tc, oracle_all_15_k_estimate_s=tc * 8, max_abs_diff_paths=float(max(np.abs(a[k] - b[k]).max() for k in ks)),
                      rel_err_vs_oracle_k15=err)), flush=True)
