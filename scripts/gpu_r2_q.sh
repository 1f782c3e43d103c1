#!/bin/bash
# round 2, GPU call Q: NaN mode with the per-feature masked denominators accumulated inside the one-pass kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_nipals.py tests/test_gpu_baseline_shapes.py tests/test_gpu_edges.py -x -q -k "nan or NaN or sparse or onepass or one_pass" > gpurun_out/q_pytest_nan.log 2>&1
tail -3 gpurun_out/q_pytest_nan.log
timeout 600 python scripts/bench_onepass.py 1.0 nan "v=one-pass trip+deflate" > gpurun_out/q_nan.json 2>&1
tail -1 gpurun_out/q_nan.json | cut -c1-900
