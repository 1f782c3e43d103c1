#!/bin/bash
# round 2, GPU call O (fixed fused standardise, NaN side kernels): fused standardise + first trip (tests, timing), whole suite, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_onepass.py -x -q > gpurun_out/o_pytest_onepass.log 2>&1
tail -5 gpurun_out/o_pytest_onepass.log
timeout 400 python scripts/bench_onepass.py 1.0 dense "v=one-pass trip+deflate" > gpurun_out/o_dense.json 2>&1
tail -1 gpurun_out/o_dense.json | cut -c1-600
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/o_pytest.log 2>&1
tail -3 gpurun_out/o_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-configs --no-cpu --verbose > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/o_bench.json").read().strip().splitlines()[-1])
print("fit ms", d["ms_per_step"], d["step_ms"], {k: round(v["ms"], 3) for k, v in d["roofline"]["per_kernel"].items()})
print("roofline", d["roofline"]["frac"], "passes", d["passes_over_X_per_step"], "actual frac", d["frac_of_hbm_peak_actual_traffic"])
print("e2e", d["e2e"]); print("nan", d["variants"]["nan_10pct"]["fit_s"]); print("parity", d["parity"]["ok"], d["parity"]["max_rel_err"])
PY
grep "e2e phases" gpurun_out/o_bench.err | tail -1
