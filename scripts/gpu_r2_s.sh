#!/bin/bash
# round 2, GPU call S: exact division by a shared divisor (reciprocal + FMA correction) in the superlevel step and the
# standardisation passes; split sums of the exchange kernel shared by several warps per item
mkdir -p gpurun_out
for a in "10000 4 1 148" "2000 8 10 592" "2000 4 10 592" "1000 2 3 1184" "640 3 2 1184"; do scripts/probes/xchg_probe $a; done > gpurun_out/s_xchg_probe.log 2>&1
cat gpurun_out/s_xchg_probe.log
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/s_pytest.log 2>&1
tail -3 gpurun_out/s_pytest.log
timeout 600 python scripts/bench_onepass.py 1.0 dense nan "v=one-pass trip+deflate" > gpurun_out/s_onepass.json 2>&1
tail -2 gpurun_out/s_onepass.json | cut -c1-900
timeout 300 python scripts/timeline.py 0.125 1 10000 > gpurun_out/s_timeline_s0125.log 2>&1
head -12 gpurun_out/s_timeline_s0125.log | cut -c1-150
