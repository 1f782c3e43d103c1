#!/bin/bash
# round 2, GPU call Y: whole suite + the full default bench line (headline, parity gate, NaN variant, e2e, other configs, CPU arm)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/y_pytest.log 2>&1
tail -3 gpurun_out/y_pytest.log
timeout 300 python scripts/prof_predict.py 2>&1 | tail -4
timeout 1500 python bench.py --verbose > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/y_bench.json").read().strip().splitlines()[-1])
print("fit ms", d["ms_per_step"], d["step_ms"], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
print("roofline", d["roofline"]["frac"], "passes", d["passes_over_X_per_step"], "actual frac", d["frac_of_hbm_peak_actual_traffic"])
print("e2e", d["e2e"]); print("nan", d["variants"]["nan_10pct"]["fit_s"]); print("parity", d["parity"]["ok"], d["parity"]["max_rel_err"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("workload", "trips_per_component", "note")} if isinstance(v, dict) else v)
print("cpu", d["cpu_baseline"])
PY
grep "e2e phases" gpurun_out/y_bench.err | tail -1
