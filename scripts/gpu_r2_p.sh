#!/bin/bash
# round 2, GPU call P: suite with the fused standardise as opt-in; kernel timelines (8-GPU-sized shard, C3-like PLS2 shard)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest.log 2>&1
tail -3 gpurun_out/p_pytest.log
timeout 300 python scripts/timeline.py 0.125 1 10000 > gpurun_out/p_timeline_s0125.log 2>&1
tail -45 gpurun_out/p_timeline_s0125.log
timeout 300 python scripts/timeline.py 0.3 10 2000 > gpurun_out/p_timeline_c3.log 2>&1
tail -45 gpurun_out/p_timeline_c3.log
