#!/bin/bash
# round 2, GPU call AC (2 GPUs): smoke(), the 2-GPU parity test and the other BASELINE configs sharded over two GPUs after the tall-shape changes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu --no-e2e --no-nan-variant > gpurun_out/ac_bench2.json 2> gpurun_out/ac_bench2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ac_bench2.json").read().strip().splitlines()[-1])
print("fit ms", round(d["ms_per_step"], 2), [round(x, 1) for x in d["step_ms"]], "parity", d["parity"]["ok"], d["parity"]["max_rel_err"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("fit_s", "predict_s", "ms_per_trip", "frac_of_measured_dgemm_whole_fit", "frac_of_hbm_peak", "error")} if isinstance(v, dict) else v)
PY
