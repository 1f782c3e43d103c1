"""Phase timing of a tall predict (C5 shape): the product kernel, the partial reduction, the device transpose and the
device->host copy.    python scripts/prof_predict.py [n] [p] [q]"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbpls_b200 import engine as E
from mbpls_b200._cabi import call
from mbpls_b200.engine import ptr, stream_ptr

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
q = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
ld = E.round_ld(n)
Xt = torch.randn((p, ld), dtype=torch.float64, device=dev)
beta = torch.randn((q, p), dtype=torch.float64, device=dev)
mean = torch.randn(p, dtype=torch.float64, device=dev)
scale = torch.rand(p, dtype=torch.float64, device=dev) + 0.5
flag = torch.zeros(1, dtype=torch.int32, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


for rep in range(3):
    torch.cuda.synchronize()
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record()
    Yh = E.skinny_gemm(Xt, n, beta, [0, p // 2, p], None, mean, scale, flag)
    e[1].record()
    Yc = Yh[:, :n].t().contiguous()
    e[2].record()
    out = E.to_host(Yh[:, :n], transpose=True)
    e[3].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"rep {rep}: skinny_gemm+reduce {e[0].elapsed_time(e[1]):.3f} ms ({8e-6 * n * p / e[0].elapsed_time(e[1]):.0f} GB/s), "
          f"transpose {e[1].elapsed_time(e[2]):.3f} ms, to_host {e[2].elapsed_time(e[3]):.3f} ms, wall {1e3 * (t1 - t0):.3f} ms", flush=True)
