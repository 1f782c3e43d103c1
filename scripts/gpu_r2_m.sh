#!/bin/bash
# round 2, GPU call M (2 GPUs): whole suite incl. the 2-GPU tests (fast epilogue, fold-parallel CV, KERNEL prefix models),
# N=2 headline, ncu of the NaN-mode one-pass kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/m_pytest.log 2>&1
tail -5 gpurun_out/m_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-parity --no-cpu --no-nan-variant > gpurun_out/m_bench2.json 2> gpurun_out/m_bench2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/m_bench2.json").read().strip().splitlines()[-1])
print("N=2 fit ms", d["ms_per_step"], d["step_ms"], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
c = d.get("configs") or {}
for k in c:
    if isinstance(c[k], dict): print("  ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in c[k].items() if kk in ("fit_s", "predict_s", "ms_per_trip", "trips_total")})
PY
CUDA_VISIBLE_DEVICES=0 timeout 900 ncu --set full --clock-control none -k regex:"fused_|masked_" -s 6 -c 6 -o gpurun_out/m_prof_nan python scripts/prof_onepass.py 10000 200000 0.1 > gpurun_out/m_ncu_nan.log 2>&1
tail -1 gpurun_out/m_ncu_nan.log
