#!/bin/bash
# round 2, GPU call F: one-kernel small fits + batched CV, predict phases, C1 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small.py tests/test_gpu_cv.py -x -q > gpurun_out/f_pytest_small.log 2>&1
tail -15 gpurun_out/f_pytest_small.log
timeout 300 python scripts/prof_predict.py > gpurun_out/f_predict.log 2>&1; cat gpurun_out/f_predict.log
MBPLS_SMALL_PATH=1 timeout 300 python scripts/bench_configs.py c1 > gpurun_out/f_c1_small.log 2>&1; cat gpurun_out/f_c1_small.log
MBPLS_SMALL_PATH=0 timeout 300 python scripts/bench_configs.py c1 > gpurun_out/f_c1_stream.log 2>&1; cat gpurun_out/f_c1_stream.log
timeout 600 python scripts/bench_cv.py > gpurun_out/f_cv.log 2>&1; cat gpurun_out/f_cv.log
timeout 900 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_methods.py -x -q > gpurun_out/f_pytest_more.log 2>&1
tail -4 gpurun_out/f_pytest_more.log
