#!/bin/bash
# round 2, GPU call V (2 GPUs): multi-GPU parity test + bench on two 8-GPU-sized shards (peer-memory exchange with the multi-CTA superlevel step)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/v_pytest_multi.log 2>&1
tail -3 gpurun_out/v_pytest_multi.log
for cfg in "A=1" "MBPLS_XCHG_MC=0"; do
  echo "== $cfg"
  env $cfg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --scale 0.25 --no-cpu --no-e2e --no-configs --no-nan-variant > gpurun_out/v_bench2.json 2> gpurun_out/v_bench2.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/v_bench2.json").read().strip().splitlines()[-1])
print("fit ms", d["ms_per_step"], d["step_ms"], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()}, d["exchange"], "parity", d["parity"]["ok"], d["parity"]["max_rel_err"])
PY
done
