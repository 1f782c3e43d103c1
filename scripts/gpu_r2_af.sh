#!/bin/bash
# round 2, GPU call AF (last): split count of the symmetric cross product (A/B), the GPU suite and the other BASELINE configs with the new rule
mkdir -p gpurun_out
timeout 60 python scripts/ab_crossprod_splits.py 2>&1 | tail -5 | tee gpurun_out/af_ab.log
timeout 110 python -m pytest tests -x -q -m gpu > gpurun_out/af_pytest.log 2>&1
tail -1 gpurun_out/af_pytest.log
timeout 70 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-nan-variant --no-parity > gpurun_out/af_bench.json 2> gpurun_out/af_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/af_bench.json").read().strip().splitlines()[-1])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("fit_s", "predict_s", "frac_of_measured_dgemm_whole_fit", "error")} if isinstance(v, dict) else v)
PY
