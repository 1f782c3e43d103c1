"""Kernel timeline of one device-resident headline fit (torch.profiler / CUPTI): where the time between the big kernels goes.

    python scripts/timeline.py <scale> [q] [n]                      (1 GPU)
    torchrun --nproc-per-node 2 scripts/timeline.py <scale> ...     (feature-sharded)

Writes gpurun_out/timeline_<tag>.json (kernel list with start / duration, runtime-API calls) and prints a digest:
total kernel time per kernel name, the idle time of the GPU between consecutive kernels grouped by (previous -> next), and the
longest gaps.  Not a benchmark: CUPTI adds a few microseconds per launch.
"""
import json
import os
import sys
import warnings
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from mbpls_b200 import MBPLS, synth
from mbpls_b200 import engine as E

SIZES_FULL = [100_000, 200_000, 300_000, 400_000]


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
    q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
    K = 20
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    sizes = [int(s * scale) for s in SIZES_FULL]
    shard = E.ShardMap.build(sizes, rank, world)
    ld = E.round_ld(n)
    Xbuf = torch.empty((max(shard.p_local, 1), ld), dtype=torch.float64, device=dev)
    Yd = synth.response(n, q, K, dev, 20261017, decay=0.85)

    def fit():
        synth.fill_feature_major(Xbuf, n, shard.lo, shard.hi, K, 20261017, noise=0.02, decay=0.85, nan_frac=0.0)
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize(dev)
        blocks = [Xbuf[shard.block_off[b]:shard.block_off[b + 1], :n].t() for b in range(len(sizes))]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = MBPLS(n_components=K, method="NIPALS", copy=False)
            m.set_runtime(device=dev, group=group, materialize=False, global_sizes=sizes, max_iter=500)
            m.fit(blocks, Yd)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1), m

    for _ in range(3):
        ms, m = fit()
    if rank == 0:
        print("untraced fit ms", round(ms, 3), "trips", sum(m.n_iter_))
    nfits = int(os.environ.get("TIMELINE_FITS", "1"))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(nfits):
            ms, m = fit()
            if rank == 0:
                print("  traced fit ms", round(ms, 3))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    print("traced fit ms", round(ms, 3))
    kern, api = [], []
    for ev in prof.events():
        name = ev.name
        tr = ev.time_range
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            kern.append((tr.start, tr.end - tr.start, name.replace("(anonymous namespace)::", "").split("(")[0][-80:]))
        elif name.startswith(("cuda", "cu")):
            api.append((tr.start, tr.end - tr.start, name))
    kern.sort()
    tag = f"s{scale}_q{q}_n{n}_w{world}"
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"fit_ms": ms, "kernels": kern, "api": sorted(api)}, open(f"gpurun_out/timeline_{tag}.json", "w"))
    tot, cnt = defaultdict(float), defaultdict(int)
    gaps, gcnt = defaultdict(float), defaultdict(int)
    longest = []
    for i, (s, d, nm) in enumerate(kern):
        tot[nm] += d
        cnt[nm] += 1
        if i:
            ps, pd, pn = kern[i - 1]
            g = s - (ps + pd)
            if g > 0:
                gaps[(pn, nm)] += g
                gcnt[(pn, nm)] += 1
                longest.append((g, pn, nm))
    span = kern[-1][0] + kern[-1][1] - kern[0][0]
    print(f"kernel span {span / 1e3:.3f} ms, busy {sum(tot.values()) / 1e3:.3f} ms, idle {(span - sum(tot.values())) / 1e3:.3f} ms")
    for nm, t in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {t / 1e3:9.3f} ms {cnt[nm]:5d} x {t / cnt[nm]:9.1f} us  {nm}")
    print("idle by (previous -> next):")
    for key, t in sorted(gaps.items(), key=lambda kv: -kv[1])[:25]:
        print(f"  {t / 1e3:9.3f} ms {gcnt[key]:5d} x {t / gcnt[key]:8.1f} us  {key[0][-40:]} -> {key[1][-40:]}")
    print("longest kernels (us):", [(round(d, 1), nm[-28:], round((s_ - kern[0][0]) / 1e3, 2)) for s_, d, nm in sorted(kern, key=lambda k: -k[1])[:14]])
    print("longest gaps (us):", [(round(g, 1), a[-24:], b[-24:]) for g, a, b in sorted(longest, reverse=True)[:12]])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
