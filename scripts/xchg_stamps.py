"""Phase timestamps of xchg_epilogue_kernel INSIDE a real fit (the probe build of the library, -DMBPLS_XCHG_STAMPS):

    nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMBPLS_XCHG_STAMPS \
         -o scripts/probes/libmbpls_b200_stamps.so mbpls_b200/csrc/*.cu
    python scripts/xchg_stamps.py <scale> <q> <n>

The stamps are those of the last launch of each fit (the last trip of the last component)."""
import ctypes as C
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mbpls_b200 import _cabi
_cabi.LIB_PATH = os.path.join(ROOT, "scripts", "probes", "libmbpls_b200_stamps.so")
from mbpls_b200 import MBPLS, synth
from mbpls_b200 import engine as E

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
K = 6
dev = torch.device("cuda:0")
sizes = [int(s * scale) for s in (100_000, 200_000, 300_000, 400_000)]
p = sum(sizes)
ld = E.round_ld(n)
Xbuf = torch.empty((p, ld), dtype=torch.float64, device=dev)
Y = synth.response(n, q, K, dev, 7, decay=0.85)
off = [0]
for s_ in sizes:
    off.append(off[-1] + s_)
lib = _cabi.load()
lib.mbpls_debug_xchg_stamps.argtypes = [C.c_void_p, C.c_int]
names = {1: "A split sums", 3: "(arrive)", 4: "last-CTA hand-off", 5: "epi 1 block scores", 6: "epi 2-3 ts, diff", 7: "epi 4 v", 9: "epi 5 u"}
acc = {k: 0.0 for k in names}
reps = 6
for rep in range(reps + 1):
    synth.fill_feature_major(Xbuf, n, 0, p, K, 8, noise=0.02, decay=0.85, nan_frac=0.0)
    blocks = [Xbuf[off[b]:off[b + 1], :n].t() for b in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = MBPLS(n_components=K, copy=False).set_runtime(materialize=False, max_iter=60).fit(blocks, Y)
    st = (C.c_ulonglong * 16)()
    lib.mbpls_debug_xchg_stamps(st, 1)
    if rep == 0:
        continue
    prev = st[0]
    for k in (1, 3, 4, 5, 6, 7, 9):
        acc[k] += (st[k] - prev) / 1e3 / reps
        prev = st[k]
print(f"n={n} p={p} q={q} trips={sum(m.n_iter_)}: last xchg launch of a fit, phases (us): " + " | ".join(f"{names[k]} {acc[k]:.1f}" for k in acc)
      + f" | total {sum(acc.values()):.1f}")
