#!/bin/bash
# round 2, GPU call K: tall kernel v2, whole suite, ncu launch list of the headline fit (full size)
mkdir -p gpurun_out
timeout 300 python scripts/prof_predict.py > gpurun_out/k_predict_tall.log 2>&1; tail -2 gpurun_out/k_predict_tall.log
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/k_pytest.log 2>&1
tail -4 gpurun_out/k_pytest.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_|xchg|standardize|xtu_kernel|record_component|segsum|begin_component|gram_partial|reduce_chunks|small_pinv|right_multiply|rows_sumsq|feature_sumsq|block_sumsq" -c 400 --csv --log-file gpurun_out/k_launches_fullsize.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-configs --no-parity --no-cpu --no-nan-variant > gpurun_out/k_ncu_bench.log 2>&1
wc -l gpurun_out/k_launches_fullsize.csv
