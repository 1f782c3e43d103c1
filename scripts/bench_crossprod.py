"""Time the FP64 tensor-core cross-product kernel against cuBLAS DGEMM (torch.matmul) on the two KERNEL shapes."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbpls_b200 import _cabi
if os.environ.get("MBPLS_LIB"):  # A/B builds of the library (scripts/probes/*.so)
    _cabi.LIB_PATH = os.path.abspath(os.environ["MBPLS_LIB"])
from mbpls_b200 import crossmethods as CM, engine as E
dev = torch.device("cuda:0")

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, p, n, kmajor in (("C5 X'X p=2000 n=1e6", 2000, 1_000_000, True), ("C2 XX' n=5000 p=50000", 50_000, 5000, False),
                           ("square X'X p=8192 n=8192", 8192, 8192, True)):
    Xt = torch.randn((p, E.round_ld(n)), dtype=torch.float64, device=dev)
    ld = Xt.shape[1]
    if kmajor:
        M, Kd, ldc = p, n, (p + 15) // 16 * 16
    else:
        M, Kd, ldc = n, p, ld
    ms = t(lambda: CM.crossprod(Xt, Xt, M, M, Kd, kmajor, ldc))
    flops = 2.0 * M * M * Kd
    A = Xt[:, :n]
    ms_blas = t(lambda: (A @ A.t()) if kmajor else (A.t() @ A))
    print(json.dumps(dict(case=name, ours_ms=ms, ours_tflops=flops / ms / 1e9, cublas_ms=ms_blas, cublas_tflops=flops / ms_blas / 1e9)), flush=True)
    del Xt, A
