"""A/B of the split-K count of the symmetric cross product: the rectangular rule (mbpls_crossprod_splits) against whole rounds of
the CTAs on / above the diagonal (mbpls_crossprod_splits_syrk), on the KERNEL / UNIPALS shapes.
    python scripts/ab_crossprod_splits.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbpls_b200 import crossmethods as CM, engine as E
from mbpls_b200._cabi import call
dev = torch.device("cuda:0")


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, p, n, kmajor in (("C5 X'X p=2000 n=1e6", 2000, 1_000_000, True), ("C5 / 8 GPUs X'X p=2000 n=125000", 2000, 125_000, True),
                           ("C2 XX' n=5000 p=50000", 50_000, 5000, False), ("square X'X p=8192 n=8192", 8192, 8192, True)):
    Xt = torch.randn((p, E.round_ld(n)), dtype=torch.float64, device=dev)
    M, Kd, ldc = (p, n, (p + 15) // 16 * 16) if kmajor else (n, p, Xt.shape[1])
    out = {"case": name}
    res = {}
    for mode in ("rectangular", "balanced"):
        os.environ["MBPLS_XP_SPLITS"] = mode
        ms = t(lambda: CM.crossprod(Xt, Xt, M, M, Kd, kmajor, ldc))
        res[mode] = CM.crossprod(Xt, Xt, M, M, Kd, kmajor, ldc)[:, :M]
        out[mode + "_ms"] = round(ms, 3)
        out[mode + "_tflops_executed"] = round(M * (M + 128.0) * Kd / ms / 1e9, 2)  # SYRK: tiles on / above the diagonal
    out["splits"] = [call("mbpls_crossprod_splits", M, M, Kd), call("mbpls_crossprod_splits_syrk", M, Kd)]
    out["max_rel_diff"] = float((res["balanced"] - res["rectangular"]).abs().max() / res["rectangular"].abs().max())
    print(json.dumps(out), flush=True)
    del Xt, res
