#!/bin/bash
# round 2, GPU call I (2 GPUs): 2-GPU parity tests, full bench at N=2 (parity gate, e2e, the other configs), NCCL vs peer-memory exchange on C3 q=10
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/i_pytest_multi.log 2>&1
tail -6 gpurun_out/i_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 1500 $TR bench.py --gpus 2 --verbose > gpurun_out/i_bench2_full.json 2> gpurun_out/i_bench2_full.err
tail -c 7000 gpurun_out/i_bench2_full.json; grep -E "e2e|Error|error|Traceback" gpurun_out/i_bench2_full.err | tail -8 | cut -c1-600
MBPLS_XCHG=nccl timeout 900 $TR bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --no-parity --no-cpu --no-nan-variant --verbose > gpurun_out/i_bench2_nccl.json 2> gpurun_out/i_bench2_nccl.err
python - <<'PY'
import json
for f in ("gpurun_out/i_bench2_full.json", "gpurun_out/i_bench2_nccl.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        c = (d.get("configs") or {}).get("c3_pls2_q10_nipals", {})
        print(f, d.get("exchange"), "fit ms", d.get("ms_per_step"), "C3 q10 fit_s", c.get("fit_s"), "ms/trip", c.get("ms_per_trip"), "trips", c.get("trips_total"))
    except Exception as e:
        print(f, "unreadable", e)
PY
