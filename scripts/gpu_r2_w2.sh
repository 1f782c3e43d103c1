#!/bin/bash
# 2 GPUs: per-step times after releasing the previous model before each timed fit
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --scale 0.25 --no-cpu --no-e2e --no-configs --no-nan-variant --no-parity > gpurun_out/w2_bench2.json 2> gpurun_out/w2_bench2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/w2_bench2.json").read().strip().splitlines()[-1])
print("fit ms", round(d["ms_per_step"], 2), [round(x, 1) for x in d["step_ms"]], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
PY
