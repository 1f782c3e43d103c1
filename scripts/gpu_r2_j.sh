#!/bin/bash
# round 2, GPU call J: tall predict kernel (tests + phases), whole suite, bench headline quick
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/j_pytest.log 2>&1
tail -5 gpurun_out/j_pytest.log
timeout 300 python scripts/prof_predict.py > gpurun_out/j_predict_tall.log 2>&1; cat gpurun_out/j_predict_tall.log
MBPLS_TALL=0 timeout 300 python scripts/prof_predict.py > gpurun_out/j_predict_old.log 2>&1; tail -1 gpurun_out/j_predict_old.log
timeout 600 ncu --set full --clock-control none -k regex:skinny_tall -s 2 -c 1 -o gpurun_out/j_prof_tall python scripts/prof_predict.py > gpurun_out/j_ncu.log 2>&1
tail -1 gpurun_out/j_ncu.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-parity --no-cpu --no-nan-variant --verbose > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/j_bench.json").read().strip().splitlines()[-1])
print("fit ms", d["ms_per_step"], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
c = d["configs"]
for k in c:
    if isinstance(c[k], dict): print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in c[k].items() if kk in ("fit_s", "predict_s", "value", "ms_per_trip", "frac_of_hbm_peak", "frac_of_measured_dgemm_whole_fit", "frac_of_hbm_peak_actual_traffic", "x_passes_per_component_equivalent")})
PY
