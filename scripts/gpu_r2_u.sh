#!/bin/bash
# round 2, GPU call U: flag-based grid sums; A/B of the multi-CTA superlevel step and of the speculative close on a PLS2 fit
mkdir -p gpurun_out
python scripts/xchg_stamps.py 0.125 1 10000 2>&1 | tail -1
python scripts/xchg_stamps.py 0.3 10 2000 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_nipals.py tests/test_gpu_baseline_shapes.py tests/test_gpu_onepass.py tests/test_gpu_edges.py -x -q > gpurun_out/u_pytest.log 2>&1
tail -2 gpurun_out/u_pytest.log
for cfg in "A=1" "MBPLS_XCHG_MC=0" "MBPLS_SPECULATE=0" "A=2"; do
  echo "== $cfg"
  env $cfg timeout 600 python scripts/bench_onepass.py 0.5 c3q10 "v=one-pass trip+deflate" 2>&1 | tail -1 | cut -c1-420
done
for cfg in "A=1" "MBPLS_XCHG_MC=0" "MBPLS_SPECULATE=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python scripts/bench_onepass.py 0.125 dense "v=one-pass trip+deflate" 2>&1 | tail -1 | cut -c1-420
done
