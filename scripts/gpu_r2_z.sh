#!/bin/bash
# round 2, GPU call Z (8 GPUs): the driver's scaling command at N = 8 after the speculative close / exchange-kernel changes
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/z_bench8.json 2> gpurun_out/z_bench8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/z_bench8.json").read().strip().splitlines()[-1])
print("fit ms", d["ms_per_step"], d["step_ms"], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["exchange"])
print("e2e", d["e2e"]["fit_s"], "nan", d["variants"]["nan_10pct"]["fit_s"], "parity", d["parity"]["ok"], d["parity"]["max_rel_err"], "skew", d["rank_skew"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("workload", "trips_per_component", "note")} if isinstance(v, dict) else v)
PY
