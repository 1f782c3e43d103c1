#!/bin/bash
# round 2, GPU call T: superlevel step spread over all CTAs of the exchange kernel (grid-wide sums)
mkdir -p gpurun_out
python scripts/xchg_stamps.py 0.125 1 10000 2>&1 | tail -1
python scripts/xchg_stamps.py 0.3 10 2000 2>&1 | tail -1
MBPLS_XCHG_MC=0 python scripts/xchg_stamps.py 0.3 10 2000 2>&1 | tail -1
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/t_pytest.log 2>&1
tail -3 gpurun_out/t_pytest.log
timeout 300 python scripts/timeline.py 0.125 1 10000 > gpurun_out/t_timeline_s0125.log 2>&1
head -12 gpurun_out/t_timeline_s0125.log | cut -c1-150
timeout 300 python scripts/timeline.py 0.3 10 2000 > gpurun_out/t_timeline_c3.log 2>&1
head -8 gpurun_out/t_timeline_c3.log | cut -c1-150
