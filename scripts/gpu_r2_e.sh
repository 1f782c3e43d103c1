#!/bin/bash
# round 2, GPU call E (2 GPUs): full test suite incl. the 2-GPU parity tests, bench at N=2 with the in-kernel exchange and with NCCL
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/e_gpus.txt
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest.log 2>&1
tail -8 gpurun_out/e_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-configs --no-parity --no-cpu --no-nan-variant --verbose > gpurun_out/e_bench2_xchg.json 2> gpurun_out/e_bench2_xchg.err
tail -c 1200 gpurun_out/e_bench2_xchg.json
MBPLS_XCHG=nccl timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-e2e --no-configs --no-parity --no-cpu --no-nan-variant --verbose > gpurun_out/e_bench2_nccl.json 2> gpurun_out/e_bench2_nccl.err
tail -c 1200 gpurun_out/e_bench2_nccl.json
timeout 1500 $TR bench.py --gpus 2 --verbose > gpurun_out/e_bench2_full.json 2> gpurun_out/e_bench2_full.err
tail -c 6000 gpurun_out/e_bench2_full.json; grep -E "e2e|Error|error" gpurun_out/e_bench2_full.err | tail -8 | cut -c1-600
