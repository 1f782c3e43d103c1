#!/bin/bash
# round 2, GPU call AA: final evidence on one GPU: suite, tall predict, full-size launch list, ncu of the final one-pass kernels
# (dense deflation + trip, NaN side kernels, exchange kernel), default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/aa_pytest.log 2>&1
tail -3 gpurun_out/aa_pytest.log
timeout 300 python scripts/prof_predict.py 2>&1 | tail -2
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_|xchg|standardize|xtu_kernel|record_component|segsum|begin_component|gram_partial|reduce_chunks|small_pinv|right_multiply|rows_sumsq|feature_sumsq|block_sumsq" -c 400 --csv --log-file gpurun_out/aa_launches_fullsize.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-configs --no-parity --no-cpu --no-nan-variant > gpurun_out/aa_ncu_bench.log 2>&1
wc -l gpurun_out/aa_launches_fullsize.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused_deflate|fused_trip|xchg_epilogue" -s 4 -c 4 -o gpurun_out/aa_prof_dense python scripts/prof_onepass.py 10000 200000 > gpurun_out/aa_ncu_dense.log 2>&1
tail -1 gpurun_out/aa_ncu_dense.log
timeout 900 ncu --set full --clock-control none -k regex:"masked_|standardize_regs" -s 1 -c 4 -o gpurun_out/aa_prof_nan python scripts/prof_onepass.py 10000 200000 0.1 > gpurun_out/aa_ncu_nan.log 2>&1
tail -1 gpurun_out/aa_ncu_nan.log
ls -la gpurun_out/aa_*.ncu-rep
