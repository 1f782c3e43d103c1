#!/bin/bash
# round 2, GPU call L (8 GPUs): full bench at N=8 with the peer-memory exchange, headline only with NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $TR bench.py --gpus 8 --verbose > gpurun_out/l_bench8_full.json 2> gpurun_out/l_bench8_full.err
grep -E "e2e|Error|error|Traceback|timeout" gpurun_out/l_bench8_full.err | tail -6 | cut -c1-400
MBPLS_XCHG=nccl timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-configs --no-parity --no-cpu --no-nan-variant > gpurun_out/l_bench8_nccl.json 2> gpurun_out/l_bench8_nccl.err
python - <<'PY'
import json
for f in ("gpurun_out/l_bench8_full.json", "gpurun_out/l_bench8_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("exchange"), "fit ms", d.get("ms_per_step"), d.get("step_ms"), {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
        print("  skew", d.get("rank_skew"))
        print("  e2e", d.get("e2e"))
        print("  parity", {k: (v.get("max_rel_err"), v.get("trips_equal")) for k, v in (d.get("parity") or {}).items() if isinstance(v, dict)})
        c = d.get("configs") or {}
        for k in c:
            if isinstance(c[k], dict): print("  ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in c[k].items() if kk in ("fit_s", "predict_s", "ms_per_trip", "trips_total", "error")})
        if "error" in c: print("  configs error", c["error"])
        print("  nan", (d.get("variants") or {}).get("nan_10pct", {}).get("fit_s"))
    except Exception as e:
        print(f, "unreadable", e)
PY
