#!/bin/bash
# round 2, GPU call B: stage hand-off protocols under racecheck, ncu of the cluster deflation kernel, new tests, parity gate
mkdir -p gpurun_out
for sy in 2 1; do
  MBPLS_FUSED_SYNC=$sy timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python scripts/prof_onepass.py 10000 600 > gpurun_out/b_racecheck_sync${sy}_10000.log 2>&1
  tail -2 gpurun_out/b_racecheck_sync${sy}_10000.log
done
for n in 16000 2000 640; do
  MBPLS_FUSED_SYNC=2 timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python scripts/prof_onepass.py $n 600 > gpurun_out/b_racecheck_sync2_$n.log 2>&1
  tail -2 gpurun_out/b_racecheck_sync2_$n.log
done
MBPLS_FUSED_SYNC=2 timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python scripts/prof_onepass.py 10000 600 0.1 > gpurun_out/b_racecheck_sync2_10000_nan.log 2>&1
tail -2 gpurun_out/b_racecheck_sync2_10000_nan.log
for sy in 2 0; do
  MBPLS_FUSED_SYNC=$sy timeout 400 python scripts/bench_onepass.py 1.0 dense "v=one-pass trip+deflate" > gpurun_out/b_dense_sy${sy}.json 2>&1
  tail -1 gpurun_out/b_dense_sy${sy}.json | cut -c1-600
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_deflate -c 2 -o gpurun_out/b_prof_cl_deflate python scripts/prof_onepass.py 10000 200000 > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log
timeout 1500 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_edges.py -x -q > gpurun_out/b_pytest.log 2>&1
tail -5 gpurun_out/b_pytest.log
timeout 900 python bench.py --steps 1 --warmup 1 --scale 0.1 --no-e2e --verbose > gpurun_out/b_bench_small.json 2> gpurun_out/b_bench_small.err
tail -c 3000 gpurun_out/b_bench_small.json; tail -5 gpurun_out/b_bench_small.err
