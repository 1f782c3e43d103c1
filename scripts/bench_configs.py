"""Secondary measurements for the other BASELINE.json configs (not the bench.py line): C1 README quickstart,
C2 KERNEL / SIMPLS PLS1 5,000 x 50,000, C3 8-block PLS2-shaped NIPALS (run as PLS1-safe q=1 and q=10 capped),
C4 NaN headline, C5 tall UNIPALS / KERNEL 1M x 2,000 + predict throughput.  Writes one JSON line per case."""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mbpls_b200 import MBPLS, synth, engine as E

warnings.simplefilter("ignore")
dev = torch.device("cuda:0")
which = set(sys.argv[1:]) or {"c1", "c2", "c3", "c4", "c5"}


def device_blocks(n, sizes, K, seed, nan_frac=0.0, noise=0.02, decay=0.85):
    p = sum(sizes)
    ld = E.round_ld(n)
    Xbuf = torch.empty((p, ld), dtype=torch.float64, device=dev)
    synth.fill_feature_major(Xbuf, n, 0, p, K, seed, noise=noise, decay=decay, nan_frac=nan_frac)
    off = np.concatenate(([0], np.cumsum(sizes)))
    return Xbuf, [Xbuf[off[b]:off[b + 1], :n].t() for b in range(len(sizes))]


def timed_fit(make, blocks_fn, Y, reps=2):
    best, m = None, None
    for _ in range(reps):
        Xbuf, blocks = blocks_fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = make().fit(blocks, Y)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        del Xbuf, blocks
    return best, m


def out(name, **kw):
    print(json.dumps(dict(case=name, **kw)), flush=True)


if "c1" in which:
    from oracle.cases import readme_quickstart
    X, y = readme_quickstart(0)
    MBPLS(n_components=3).fit(X, y)
    t0 = time.perf_counter()
    for _ in range(20):
        m = MBPLS(n_components=3).fit(X, y)
    out("C1 README quickstart NIPALS 40x(200+250) K=3 (host numpy in, numpy attrs out)", fit_ms=(time.perf_counter() - t0) / 20 * 1e3,
        trips=m.n_iter_)

if "c2" in which:
    n, p, K = 5000, 50000, 10
    Y = synth.response(n, 1, K, dev, 11, decay=0.85)
    for method in ("SIMPLS", "KERNEL", "UNIPALS", "NIPALS"):
        dt, m = timed_fit(lambda: MBPLS(n_components=K, method=method, copy=False, calc_all=False).set_runtime(materialize=False),
                          lambda: device_blocks(n, [p], K, 12), Y)
        out(f"C2 PLS1 {method} n=5000 p=50000 K=10 calc_all=False (device-resident input)", fit_s=dt,
            x_gbytes=8e-9 * n * p)

if "c3" in which:
    n, K = 2000, 20
    sizes = [20000, 35000, 60000, 95000, 140000, 180000, 220000, 450000]
    Y = synth.response(n, 1, K, dev, 21, decay=0.85)
    dt, m = timed_fit(lambda: MBPLS(n_components=K, copy=False).set_runtime(materialize=False, max_iter=500),
                      lambda: device_blocks(n, sizes, K, 22), Y)
    trips = m.n_iter_
    out("C3 8-block NIPALS n=2000 p=1.2M K=20 q=1", fit_s=dt, trips=trips,
        algorithmic_gbs=16.0 * n * sum(sizes) * (1 + K + sum(trips)) / dt / 1e9)
    Y10 = synth.response(n, 10, K, dev, 21, decay=0.85)
    dt, m = timed_fit(lambda: MBPLS(n_components=K, copy=False).set_runtime(materialize=False, max_iter=300),
                      lambda: device_blocks(n, sizes, K, 22), Y10, reps=1)
    trips = m.n_iter_
    out("C3 8-block NIPALS n=2000 p=1.2M K=20 q=10 (trips capped at 300/component)", fit_s=dt, trips=trips,
        algorithmic_gbs=16.0 * n * sum(sizes) * (1 + K + sum(trips)) / dt / 1e9)

if "c4" in which:
    n, K = 10000, 20
    sizes = [100000, 200000, 300000, 400000]
    Y = synth.response(n, 1, K, dev, 31, decay=0.85)
    dt, m = timed_fit(lambda: MBPLS(n_components=K, copy=False, sparse_data=True).set_runtime(materialize=False, max_iter=300),
                      lambda: device_blocks(n, sizes, K, 32, nan_frac=0.10), Y)
    trips = m.n_iter_
    out("C4 NaN NIPALS n=10000 p=1M 10% NaN K=20 q=1", fit_s=dt, trips=trips,
        algorithmic_gbs=16.0 * n * sum(sizes) * (1 + K + sum(trips)) / dt / 1e9)

if "c5" in which:
    n, K = 1_000_000, 30
    sizes = [1000, 1000]
    Y = synth.response(n, 4, K, dev, 41, decay=0.9)
    for method in ("KERNEL", "UNIPALS"):
        dt, m = timed_fit(lambda: MBPLS(n_components=K, method=method, copy=False, calc_all=False).set_runtime(materialize=False),
                          lambda: device_blocks(n, sizes, K, 42, decay=0.9), Y, reps=1)
        out(f"C5 tall {method} n=1M p=2000 K=30 q=4 calc_all=False", fit_s=dt, x_gbytes=16.0)
    Xbuf, blocks = device_blocks(n, sizes, K, 43, decay=0.9)
    m.predict(blocks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    yh = m.predict(blocks)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out("C5 predict 1M x 2000 (device-resident input, numpy output)", predict_s=dt, rows_per_s=n / dt, gbs=16.0 / dt)
