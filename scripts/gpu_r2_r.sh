#!/bin/bash
# round 2, GPU call R: components closed speculatively (predicated record + deflation behind the trips); timeline of an 8-GPU-sized shard
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r_pytest.log 2>&1
tail -3 gpurun_out/r_pytest.log
timeout 300 python scripts/timeline.py 0.125 1 10000 > gpurun_out/r_timeline_s0125.log 2>&1
head -12 gpurun_out/r_timeline_s0125.log | cut -c1-150; grep -A8 "idle by" gpurun_out/r_timeline_s0125.log | cut -c1-150
