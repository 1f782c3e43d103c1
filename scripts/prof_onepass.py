"""One small dense NIPALS fit through the one-pass kernels, for `ncu -k regex:fused` captures.
    python scripts/prof_onepass.py [n] [p] [nan_frac]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mbpls_b200 import MBPLS, synth, engine as E

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
nan_frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
dev = torch.device("cuda:0")
Xbuf = torch.empty((p, E.round_ld(n)), dtype=torch.float64, device=dev)
synth.fill_feature_major(Xbuf, n, 0, p, 3, 5, noise=0.02, decay=0.85, nan_frac=nan_frac)
Y = synth.response(n, 1, 3, dev, 31, decay=0.85)
m = MBPLS(n_components=3, copy=False, sparse_data=nan_frac > 0).set_runtime(materialize=False, one_pass=True).fit([Xbuf[:, :n].t()], Y)
torch.cuda.synchronize()
print("trips", m.n_iter_)
