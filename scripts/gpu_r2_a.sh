#!/bin/bash
# round 2, GPU call A: cluster deflation / trip kernels -- parity, A/B timing, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt
timeout 1200 python -m pytest tests/test_gpu_onepass.py -x -q > gpurun_out/a_pytest_onepass.log 2>&1
tail -5 gpurun_out/a_pytest_onepass.log
for cl in 1 0; do for sy in 0 1; do
  MBPLS_FUSED_CLUSTER=$cl MBPLS_FUSED_SYNC=$sy timeout 400 python scripts/bench_onepass.py 1.0 dense "v=one-pass trip+deflate" > gpurun_out/a_dense_cl${cl}_sy${sy}.json 2>&1
  tail -1 gpurun_out/a_dense_cl${cl}_sy${sy}.json | cut -c1-600
done; done
timeout 400 python scripts/bench_onepass.py 1.0 n20k "v=one-pass trip+deflate" "v=two-pass" > gpurun_out/a_n20k.json 2>&1
tail -4 gpurun_out/a_n20k.json | cut -c1-600
MBPLS_FUSED_CLUSTER=1 timeout 300 python scripts/bench_onepass.py 1.0 nan "v=one-pass trip+deflate" > gpurun_out/a_nan_cl1.json 2>&1
tail -1 gpurun_out/a_nan_cl1.json | cut -c1-600
for n in 10000 16000 2000; do
  timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python scripts/prof_onepass.py $n 600 > gpurun_out/a_racecheck_$n.log 2>&1
  tail -3 gpurun_out/a_racecheck_$n.log
done
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python scripts/prof_onepass.py 10000 600 0.1 > gpurun_out/a_memcheck_10000_nan.log 2>&1
tail -3 gpurun_out/a_memcheck_10000_nan.log
