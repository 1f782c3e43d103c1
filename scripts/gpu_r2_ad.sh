#!/bin/bash
# round 2, GPU call AD (last but one): the driver's own round-end sequence on the final HEAD: smoke(), the GPU suite, the default bench line,
# the reference arm on a short run
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/ad_pytest.log 2>&1
tail -1 gpurun_out/ad_pytest.log
timeout 400 python bench.py > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ad_bench.json").read().strip().splitlines()[-1])
print("fit ms", d["ms_per_step"], d["step_ms"], "roofline", d["roofline"]["frac"], "e2e", d["e2e"]["fit_s"], "parity", d["parity"]["ok"],
      d["parity"]["max_rel_err"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("cpu", d["cpu_baseline"])
PY
