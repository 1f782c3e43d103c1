#!/bin/bash
# round 2, GPU call D: full single-GPU test suite, bench with the other BASELINE configs (small scale first, then full)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/d_pytest.log 2>&1
tail -6 gpurun_out/d_pytest.log
timeout 600 python bench.py --steps 1 --warmup 1 --scale 0.05 --no-cpu --no-parity --verbose > gpurun_out/d_bench_small.json 2> gpurun_out/d_bench_small.err
tail -c 2500 gpurun_out/d_bench_small.json; tail -4 gpurun_out/d_bench_small.err | cut -c1-1500
timeout 1500 python bench.py --verbose > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -c 4000 gpurun_out/d_bench.json; grep -E "e2e|configs" gpurun_out/d_bench.err | cut -c1-3000
