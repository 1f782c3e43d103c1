"""Kernel timeline (torch.profiler) of one SIMPLS / KERNEL / UNIPALS fit at the C2 shape (PLS1, n = 5,000 x p = 50,000, K = 10):
GPU busy time per kernel name and idle time between kernels -- is the fit bound by the host's launch rate?
    python scripts/timeline_methods.py [method] [n] [p] [K] [q] [blocks] [calc_all]"""
import os
import sys
import warnings
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from mbpls_b200 import MBPLS, synth
from mbpls_b200 import engine as E

method = sys.argv[1] if len(sys.argv) > 1 else "SIMPLS"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
p = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
q = int(sys.argv[5]) if len(sys.argv) > 5 else 1
nblocks = int(sys.argv[6]) if len(sys.argv) > 6 else 1
calc_all = (sys.argv[7] != "0") if len(sys.argv) > 7 else True
dev = torch.device("cuda:0")
ld = E.round_ld(n)
Xbuf = torch.empty((p, ld), dtype=torch.float64, device=dev)
Y = synth.response(n, q, K, dev, 5, decay=0.85)
bounds = [p * b // nblocks for b in range(nblocks + 1)]


def fit():
    synth.fill_feature_major(Xbuf, n, 0, p, K, 6, noise=0.02, decay=0.85, nan_frac=0.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MBPLS(n_components=K, method=method, copy=False, calc_all=calc_all).set_runtime(materialize=False).fit([Xbuf[bounds[b]:bounds[b + 1], :n].t() for b in range(nblocks)], Y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), m


for _ in range(3):
    ms, m = fit()
print(method, "untraced fit ms", round(ms, 3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    synth.fill_feature_major(Xbuf, n, 0, p, K, 6, noise=0.02, decay=0.85, nan_frac=0.0)
    torch.cuda.synchronize()
    mark0 = torch.cuda.Event(enable_timing=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MBPLS(n_components=K, method=method, copy=False, calc_all=calc_all).set_runtime(materialize=False).fit([Xbuf[bounds[b]:bounds[b + 1], :n].t() for b in range(nblocks)], Y)
    torch.cuda.synchronize()
kern = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tr = ev.time_range
        kern.append((tr.start, tr.end - tr.start, ev.name.replace("(anonymous namespace)::", "").split("(")[0][-60:]))
kern.sort()
first = min(i for i, k in enumerate(kern) if "standardize" in k[2])  # the fit starts with the standardisation of Y
kern = kern[first:]
tot, cnt = defaultdict(float), defaultdict(int)
for s, d, nm in kern:
    tot[nm] += d
    cnt[nm] += 1
span = kern[-1][0] + kern[-1][1] - kern[0][0]
busy = sum(tot.values())
print(f"{len(kern)} kernels / copies, span {span / 1e3:.3f} ms, busy {busy / 1e3:.3f} ms, idle {(span - busy) / 1e3:.3f} ms")
for nm, t in sorted(tot.items(), key=lambda kv: -kv[1])[:26]:
    print(f"  {t / 1e3:8.3f} ms {cnt[nm]:5d} x {t / cnt[nm]:8.1f} us  {nm}")
