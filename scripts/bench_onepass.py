"""One-pass vs two-pass NIPALS kernels, per-kernel CUDA-event times (device-resident input, one process).

    python scripts/bench_onepass.py [scale] [cases...]     cases: dense nan c3 c3q10 mid
Prints one JSON line per (case, variant): fit seconds, trips, and ms / GB/s per kernel class."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mbpls_b200 import MBPLS, synth, engine as E

warnings.simplefilter("ignore")
dev = torch.device("cuda:0")
args = sys.argv[1:]
scale = float(args[0]) if args and args[0].replace(".", "").isdigit() else 1.0
only = [a[2:] for a in args if a.startswith("v=")]  # v=<variant name>: run only these variants
args = [a for a in args if not a.startswith("v=")]
which = set(a for a in args if not a.replace(".", "").isdigit()) or {"dense", "nan", "c3"}

VARIANTS = (("two-pass", dict(one_pass=False)), ("one-pass trip", dict(one_pass=True, one_pass_deflate=False)),
            ("one-pass trip+deflate", dict(one_pass=True)))
if only:
    VARIANTS = tuple(v for v in VARIANTS if v[0] in only)
MULT = {"colden": 1.0 / 64, "rowden": 1.0 / 64, "trip": 1.0, "xtu": 1.0, "xw": 1.0, "deflate": 2.0, "loadings": 1.0, "standardize": 2.0}


def run(case, n, sizes, K, q, nan_frac, max_iter=300, variants=VARIANTS, reps=2):
    p = sum(sizes)
    ld = E.round_ld(n)
    Xbuf = torch.empty((p, ld), dtype=torch.float64, device=dev)
    Y = synth.response(n, q, K, dev, 31, decay=0.85)
    off = np.concatenate(([0], np.cumsum(sizes)))
    for name, rt in variants:
        best = None
        for _ in range(reps):
            synth.fill_feature_major(Xbuf, n, 0, p, K, 32, noise=0.02, decay=0.85, nan_frac=nan_frac)
            blocks = [Xbuf[off[b]:off[b + 1], :n].t() for b in range(len(sizes))]
            prof = {}
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m = MBPLS(n_components=K, copy=False, sparse_data=nan_frac > 0).set_runtime(materialize=False, max_iter=max_iter,
                                                                                        profile=prof, **rt).fit(blocks, Y)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, prof, list(m.n_iter_))
        dt, prof, trips = best
        kern = {}
        for key, ev in prof.items():
            ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            kern[key] = dict(ms=round(ms, 3), launches=len(ev), gbs=round(MULT.get(key, 1.0) * 8.0 * n * p / ms / 1e6, 1))
        print(json.dumps(dict(case=case, variant=name, n=n, p=p, K=K, q=q, nan_frac=nan_frac, fit_s=round(dt, 4),
                              trips_total=sum(trips), canonical_gbs=round(16.0 * n * p * (1 + K + sum(trips)) / dt / 1e9, 1),
                              kernels=kern)), flush=True)
    del Xbuf


S4 = [int(s * scale) for s in (100_000, 200_000, 300_000, 400_000)]
if "dense" in which:
    run("C4-dense headline", 10_000, S4, 20, 1, 0.0)
if "nan" in which:
    run("C4 10% NaN", 10_000, S4, 20, 1, 0.10)
if "c3" in which:
    S8 = [int(s * scale) for s in (20000, 35000, 60000, 95000, 140000, 180000, 220000, 450000)]
    run("C3 8 blocks n=2000 q=1", 2000, S8, 20, 1, 0.0)
if "c3q10" in which:
    S8 = [int(s * scale) for s in (20000, 35000, 60000, 95000, 140000, 180000, 220000, 450000)]
    run("C3 8 blocks n=2000 q=10", 2000, S8, 20, 10, 0.0, reps=1, variants=VARIANTS[::2])
if "c3nan" in which:
    S8 = [int(s * scale) for s in (20000, 35000, 60000, 95000, 140000, 180000, 220000, 450000)]
    run("C3 8 blocks n=2000 q=1 10% NaN", 2000, S8, 20, 1, 0.10)
    run("n=4000 p=500k q=1 10% NaN", 4000, [int(500_000 * scale)], 10, 1, 0.10)
if "n20k" in which:  # features split over CTA pairs for both passes
    run("n=20000 p=500k q=1", 20_000, [int(500_000 * scale)], 10, 1, 0.0)
    run("n=16000 p=500k q=1", 16_000, [int(500_000 * scale)], 10, 1, 0.0)
if "mid" in which:
    run("n=4000 p=500k q=1", 4000, [int(500_000 * scale)], 10, 1, 0.0)
    run("n=1000 p=2M q=1", 1000, [int(2_000_000 * scale)], 10, 1, 0.0)
