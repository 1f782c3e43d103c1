#!/bin/bash
# round 2, GPU call AB (2 GPUs): suite incl. the 2-GPU parity test, then per-step times of two 8-GPU-sized shards
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1
tail -3 gpurun_out/ab_pytest.log
bash scripts/gpu_r2_w2.sh
