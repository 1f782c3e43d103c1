#!/bin/bash
# round 2, GPU call W (2 GPUs): A/B of the multi-CTA superlevel step between GPUs, warm box, alternating order
mkdir -p gpurun_out
for cfg in "MBPLS_XCHG_MC=0" "A=1" "MBPLS_XCHG_MC=0" "A=1"; do
  echo "== $cfg"
  env $cfg timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --scale 0.25 --no-cpu --no-e2e --no-configs --no-nan-variant --no-parity > gpurun_out/w_bench2.json 2> gpurun_out/w_bench2.err
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/w_bench2.json").read().strip().splitlines()[-1])
print("fit ms", round(d["ms_per_step"], 2), [round(x, 1) for x in d["step_ms"]], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
PY
done
