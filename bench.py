#!/usr/bin/env python
"""bench.py -- MB-PLS NIPALS fit throughput on B200 (BASELINE.json: "MB-PLS fit s & HBM GB/s
(n=10k, p=1M, 20 LV) at 1/2/4/8 B200 vs host CPU").

One *step* = one complete fp64 multiblock NIPALS fit (column standardisation, 20 latent variables,
finalisation of R_/beta_) of n=10,000 samples x p=1,000,000 features in 4 blocks (100k/200k/300k/400k)
with a single response (PLS1: the NIPALS loop takes exactly 2 trips per component, so the work per
step is data-independent; see DESIGN.md "Benchmark workload").  Synthetic latent-structure data is
generated on the device (feature-major); because `fit` standardises and deflates X in place the input
is regenerated before every step, outside the timed region.

metric/unit: algorithmic HBM GB/s = 16*n*p*(1 + K + sum_k I_k) bytes / fit seconds (the traffic of an
exact two-pass NIPALS); `fit_s` is the absolute time.  The one-pass kernels (csrc/fused.cu) read X once per trip, so
the canonical figure can exceed the HBM peak; `hbm_gbs_actual_traffic` / `passes_over_X_per_step` report what the
kernels really move and `roofline` times the dominant kernel against its own bytes.  N > 1: the feature axis is sharded (strong scaling), one NCCL
all-reduce of the (n x B + B) partial block scores per trip.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MB-PLS NIPALS fit: algorithmic HBM GB/s (n=10k, p=1M, 4 blocks, 20 LV, fp64)"
SIZES_FULL = (100_000, 200_000, 300_000, 400_000)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=10_000)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the feature counts (debug / small GPUs)")
    ap.add_argument("--components", type=int, default=20)
    ap.add_argument("--q", type=int, default=1)
    ap.add_argument("--nan-frac", type=float, default=0.0)
    ap.add_argument("--max-iter", type=int, default=200, help="safety cap on trips per component for both arms")
    ap.add_argument("--noise", type=float, default=0.02)
    ap.add_argument("--decay", type=float, default=0.85)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-nan-variant", action="store_true", help="skip the extra 10 %% NaN fit reported under `variants`")
    ap.add_argument("--two-pass", action="store_true", help="force the two-pass NIPALS kernels (X'u and X w as separate reads)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-p", type=int, default=0, help="features of the CPU sample (0: auto)")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--verbose", action="store_true")
    return ap.parse_args()


VERBOSE = False
_T0 = time.time()


def log(*a):
    if VERBOSE:
        print(f"[bench +{time.time() - _T0:7.1f}s]", *a, file=sys.stderr, flush=True)


def fit_bytes(n, p, K, trips):
    """Canonical algorithmic traffic of one NIPALS fit (SURVEY.md 8d / BASELINE.md section 4)."""
    return 16.0 * n * p * (1 + K + sum(trips))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the numpy oracle port of mbpls/mbpls.py on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(args, p_sample, repeats=1):
    import numpy as np
    import threadpoolctl
    from oracle import OracleMBPLS
    from oracle.cases import latent_blocks
    n, K = args.n, args.components
    frac = [s / sum(SIZES_FULL) for s in SIZES_FULL]
    sizes = [max(8, int(round(p_sample * f))) for f in frac]
    X, Y = latent_blocks(n, sizes, args.q, K, seed=args.seed % 100000, noise=args.noise, decay=args.decay,
                         nan_frac=args.nan_frac)
    best, trips = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = OracleMBPLS(n_components=K, method="NIPALS", sparse_data=args.nan_frac > 0,
                            max_iter=args.max_iter).fit([x.copy() for x in X], Y.copy())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        trips = list(m.n_iter_)
    info = threadpoolctl.threadpool_info()
    threads = max([i.get("num_threads", 1) for i in info] or [1])
    p = sum(sizes)
    return dict(seconds=best, gbs=fit_bytes(n, p, K, trips) / best / 1e9, trips=trips, cores=threads,
                sample=f"oracle numpy/OpenBLAS port of mbpls.py NIPALS, n={n}, p={p} in 4 blocks, q={args.q}, K={K}, "
                       f"{threads} BLAS threads of {os.cpu_count()} cpus, best of {repeats}")


def auto_cpu_p(args, budget_s):
    """Pick the CPU sample width so one fit takes roughly `budget_s` (about 3 GB/s measured on 8-16 cores)."""
    if args.cpu_p:
        return args.cpu_p
    per_feature = 16.0 * args.n * (1 + args.components + (2 if args.q == 1 else 40) * args.components) / 3e9
    return int(max(400, min(40_000, budget_s / per_feature)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    budget = max(3.0, 150.0 / (steps + warm))
    p_sample = auto_cpu_p(args, budget)
    for _ in range(warm):
        cpu_sample(args, p_sample)
    times, res = [], None
    for _ in range(steps):
        res = cpu_sample(args, p_sample)
        times.append(res["seconds"])
    sec = sum(times) / len(times)
    p = p_sample
    gbs = res["gbs"] * res["seconds"] / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, note=f"bounded CPU sample: p={p} features instead of {int(sum(SIZES_FULL) * args.scale)}"),
        "trips_per_component": res["trips"],
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, note=None):
    sizes = [int(s * args.scale) for s in SIZES_FULL]
    cfg = {"workload": f"NIPALS MB-PLS fit n={args.n} p={sum(sizes)} blocks={sizes} q={args.q} K={args.components} "
                       f"nan_frac={args.nan_frac} standardize=True calc_all=True max_tol=1e-14; latent-structure data "
                       f"(r=K+5 factors, decay {args.decay}, noise {args.noise})",
           "parallelism": f"feature-sharded x{args.gpus}", "l2_policy": "inputs (>=10 GB per GPU) larger than the 126 MB L2",
           "timing": "per-step CUDA-event bracket (barrier + synchronize both sides); X regenerated between steps outside it"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/mbpls_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [s for s in sm if s > 0]
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mbpls_b200 import MBPLS, _cabi, synth
    from mbpls_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    n, K, q = args.n, args.components, args.q
    sizes = [int(s * args.scale) for s in SIZES_FULL]
    p = sum(sizes)
    shard = E.ShardMap.build(sizes, rank, world)
    ld = E.round_ld(n)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    Xbuf = torch.empty((max(shard.p_local, 1), ld), dtype=torch.float64, device=dev)
    Yd = synth.response(n, q, K, dev, args.seed, decay=args.decay)  # n x q, identical on every rank

    def regenerate(nan_frac=None):
        synth.fill_feature_major(Xbuf, n, shard.lo, shard.hi, K, args.seed, noise=args.noise, decay=args.decay,
                                 nan_frac=args.nan_frac if nan_frac is None else nan_frac)

    def local_blocks():
        return [Xbuf[shard.block_off[b]:shard.block_off[b + 1], :n].t() for b in range(len(sizes))]

    def barrier():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)

    one_pass = False if args.two_pass else None

    def make_model(profile=None, materialize=False, sparse=None, rec=None):
        m = MBPLS(n_components=K, method="NIPALS", standardize=True, calc_all=True,
                  sparse_data=(args.nan_frac > 0) if sparse is None else sparse, copy=False)
        m.set_runtime(device=dev, group=group, materialize=materialize, global_sizes=sizes, profile=profile,
                      max_iter=args.max_iter, one_pass=one_pass, deflate_rec=rec)
        return m

    def one_fit(profile=None, nan_frac=None, rec=None):
        regenerate(nan_frac)
        barrier()
        log("regenerated")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = make_model(profile, sparse=None if nan_frac is None else nan_frac > 0, rec=rec).fit(local_blocks(), Yd)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
        return float(ms.item()), m

    log("setup done; p_local", shard.p_local)
    for _ in range(max(args.warmup, 0)):
        ms, _m = one_fit()
        log("warmup fit ms", ms, "trips", _m.n_iter_)
    clocks = Clocks(local) if rank == 0 else None
    profile, times, model = {}, [], None
    launches0 = _cabi.launch_count
    for _ in range(max(args.steps, 1)):
        ms, model = one_fit(profile)
        times.append(ms)
        log("timed fit ms", ms)
    launches = _cabi.launch_count - launches0
    clk = clocks.stop() if clocks else None
    # BASELINE config 4 as written: the same fit with 10 % i.i.d. NaN holes (sparse_data=True), reported beside the headline
    variants = {}
    if args.nan_frac == 0 and not args.no_nan_variant:
        one_fit(nan_frac=0.10)
        nan_ms, nan_model = one_fit(nan_frac=0.10)
        nt = list(nan_model.n_iter_)
        variants["nan_10pct"] = {"fit_s": nan_ms / 1e3, "value": fit_bytes(n, p, K, nt) / (nan_ms / 1e3) / 1e9, "unit": "GB/s",
                                 "trips_per_component": nt, "kernels": "one-pass kernels with NaN read as zero + masked denominators from the NaN bit matrix (DESIGN.md 3a)"}
        del nan_model
        # opt-in deflation that carries x_j . u0 per feature instead of keeping u0 in shared memory: faster,
        # but its first-trip weights are noisier, so trip counts can exceed the reference's; reported for information only
        one_fit(rec=True)
        rec_ms, rec_model = one_fit(rec=True)
        rt_ = list(rec_model.n_iter_)
        variants["recurrence_deflation_opt_in"] = {"fit_s": rec_ms / 1e3, "value": fit_bytes(n, p, K, rt_) / (rec_ms / 1e3) / 1e9,
                                                   "unit": "GB/s", "trips_per_component": rt_}
        del rec_model
    trips = list(model.n_iter_)
    ms_step = sum(times) / len(times)
    value = fit_bytes(n, p, K, trips) / (ms_step / 1e3) / 1e9

    # X-sized transfers the kernels of one step really make (reads + writes), from the launches timed above
    passes = sum(len(profile.get(k, [])) * m_ for k, m_ in (("trip", 1), ("xtu", 1), ("xw", 1), ("deflate", 2), ("loadings", 1),
                                                             ("standardize", 2))) / max(len(times), 1)
    # per-kernel roofline (rank 0's shard): algorithmic bytes of one launch / mean CUDA-event duration
    def mean_ms(key):
        ev = profile.get(key, [])
        return sum(a.elapsed_time(b) for a, b in ev) / len(ev) if ev else None

    x_bytes_local = 8.0 * n * shard.p_local
    kern = {}
    for key, mult in (("trip", 1.0), ("xtu", 1.0), ("xw", 1.0), ("deflate", 2.0), ("loadings", 1.0), ("standardize", 2.0)):
        t = mean_ms(key)
        if t:
            kern[key] = {"ms": t, "launches": len(profile[key]), "algorithmic_bytes": mult * x_bytes_local,
                         "gbs": mult * x_bytes_local / (t / 1e3) / 1e9}
    # multi-GPU: every trip ends in an all-reduce, so the slowest rank sets the pace; report the spread of the per-rank means
    rank_skew = None
    if world > 1:
        try:
            keys = ("trip", "deflate", "standardize", "loadings")
            mine = torch.tensor([kern[k]["ms"] if k in kern else 0.0 for k in keys], dtype=torch.float64, device=dev)
            allv = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine, group=group)
            stack = torch.stack(allv).cpu()
            rank_skew = {k: {"min_ms": float(stack[:, i].min()), "max_ms": float(stack[:, i].max())} for i, k in enumerate(keys)}
        except Exception as exc:  # diagnostics must never take the benchmark line down
            rank_skew = {"error": repr(exc)[:200]}
    # dominant kernel by share of the step: the loadings+deflation(+next first trip) pass when the one-pass kernels run
    # (1 read + 1 write of X per launch), else the X w pass of the two-pass kernels
    dominant = max(kern, key=lambda k: kern[k]["ms"] * kern[k]["launches"]) if kern else None
    names = {"trip": "fused_trip_kernel (X'u and X w in one read)", "deflate": "fused_deflate_kernel / loadings_deflate_*",
             "xw": "xw_kernel", "xtu": "xtu_kernel", "standardize": "standardize_*", "loadings": "xtu_kernel (loadings)"}
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        entry = prof.get(dominant, {})
        if dominant and entry.get("algorithmic_bytes_per_launch"):
            # the capture may have been taken on a different shard size: scale by the algorithmic bytes
            traffic = entry["dram_bytes_per_launch"] * kern[dominant]["algorithmic_bytes"] / entry["algorithmic_bytes_per_launch"]
    except Exception:
        pass
    roofline = None
    if dominant:
        roofline = {"bound": "hbm", "kernel": names.get(dominant, dominant) + (" [NaN mode]" if args.nan_frac > 0 else ""),
                    "achieved": kern[dominant]["gbs"], "peak": peak_gbs, "unit": "GB/s",
                    "frac": kern[dominant]["gbs"] / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kern[dominant]["algorithmic_bytes"], "per_kernel": kern}

    # ---- end to end through the public API with HOST buffers (pinned, row-major like the reference's inputs)
    e2e = None
    del model
    if not args.no_e2e:
        import psutil
        need = 8.0 * n * shard.p_local
        avail = psutil.virtual_memory().available / max(world, 1)
        ok = torch.tensor([1 if need < 0.6 * avail else 0], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()):
            log("e2e: building pinned host copy")
            regenerate()
            host = []
            for b in range(len(sizes)):
                o0, o1 = shard.block_off[b], shard.block_off[b + 1]
                h = torch.empty((n, o1 - o0), dtype=torch.float64, pin_memory=(o1 > o0))
                for c in range(0, o1 - o0, 8192):
                    c1 = min(o1 - o0, c + 8192)
                    h[:, c:c1].copy_(Xbuf[o0 + c:o0 + c1, :n].t())
                host.append(h)
            Yh = Yd.cpu().pin_memory()
            torch.cuda.synchronize(dev)
            del Xbuf
            torch.cuda.empty_cache()
            e_times, d2h = [], 0
            for it in range(1 + max(1, args.e2e_steps)):
                barrier()
                t0 = time.perf_counter()
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    m = MBPLS(n_components=K, method="NIPALS", sparse_data=args.nan_frac > 0, copy=True)
                    m.set_runtime(device=dev, group=group, materialize=True, global_sizes=sizes, max_iter=args.max_iter)
                    m.fit(host, Yh)
                barrier()
                dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX, group=group)
                log("e2e fit s", float(dt.item()))
                if it > 0:
                    e_times.append(float(dt.item()))
                d2h = sum(a.nbytes for a in [m.Ts_, m.U_, m.V_, m.R_, m.beta_, m.A_] + m.T_ + m.W_ + m.P_ + m.W_non_normal_)
                e_trips = list(m.n_iter_)
                del m
                torch.cuda.empty_cache()
            es = sum(e_times) / len(e_times)
            e2e = {"value": fit_bytes(n, p, K, e_trips) / es / 1e9, "unit": "GB/s", "fit_s": es,
                   "h2d_bytes_per_step": int(8 * n * shard.p_local + 8 * n * q), "d2h_bytes_per_step": int(d2h),
                   "what": "MBPLS.fit(list of pinned row-major host blocks, Y) -> numpy attributes; wall clock, max over ranks"}
        else:
            e2e = {"value": None, "unit": "GB/s", "skipped": "host RAM too small for the full-size input"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        log("cpu baseline sample p", auto_cpu_p(args, 15.0))
        r = cpu_sample(args, auto_cpu_p(args, 15.0))
        log("cpu baseline done", r["seconds"])
        cpu = {"value": r["gbs"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "seconds": r["seconds"], "trips_per_component": r["trips"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": len(times), "warmup": args.warmup,
            "ms_per_step": ms_step, "fit_s": ms_step / 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (device-generated latent-structure blocks)",
            "config": workload_config(args), "trips_per_component": trips, "algorithmic_bytes_per_step": fit_bytes(n, p, K, trips),
            "frac_of_hbm_peak": value / (peak_gbs * world),
            "frac_of_hbm_peak_actual_traffic": passes * 8.0 * n * p / (ms_step / 1e3) / 1e9 / (peak_gbs * world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clk, "step_ms": times, "variants": variants, "rank_skew": rank_skew,
            "passes_over_X_per_step": passes, "hbm_gbs_actual_traffic": passes * 8.0 * n * p / (ms_step / 1e3) / 1e9,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global VERBOSE
    args = parse()
    VERBOSE = args.verbose
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
