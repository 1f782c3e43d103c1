#!/bin/bash
# round 2, GPU call X (2 GPUs): where do the sporadic stalls of the multi-CTA superlevel step between GPUs come from?
mkdir -p gpurun_out
TIMELINE_FITS=6 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scripts/timeline.py 0.25 1 10000 > gpurun_out/x_timeline2.log 2>&1
grep -v Warning gpurun_out/x_timeline2.log | cut -c1-600 | head -60
