#!/bin/bash
# round 2, GPU call G: whole single-GPU suite, predict phases after the folded scaling, C1 one-kernel timing
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest.log 2>&1
tail -8 gpurun_out/g_pytest.log
timeout 300 python scripts/prof_predict.py > gpurun_out/g_predict.log 2>&1; cat gpurun_out/g_predict.log
MBPLS_SMALL_PATH=1 timeout 300 python scripts/bench_configs.py c1 > gpurun_out/g_c1_small.log 2>&1; cat gpurun_out/g_c1_small.log
