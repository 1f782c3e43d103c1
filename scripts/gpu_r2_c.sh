#!/bin/bash
# round 2, GPU call C: mbarrier hand-off (try_wait) under racecheck, CTA-scope exchange timing, cluster-size probe,
# BASELINE-shaped parity tests, full default bench
mkdir -p gpurun_out
scripts/probes/cluster_probe > gpurun_out/c_cluster_probe.txt 2>&1; cat gpurun_out/c_cluster_probe.txt
for n in 10000 640; do
  MBPLS_FUSED_SYNC=2 timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python scripts/prof_onepass.py $n 600 > gpurun_out/c_racecheck_sync2_$n.log 2>&1
  tail -2 gpurun_out/c_racecheck_sync2_$n.log
done
for sy in 0 2; do
  MBPLS_FUSED_SYNC=$sy timeout 400 python scripts/bench_onepass.py 1.0 dense "v=one-pass trip+deflate" > gpurun_out/c_dense_sy${sy}.json 2>&1
  tail -1 gpurun_out/c_dense_sy${sy}.json | cut -c1-600
done
timeout 1500 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_nipals.py tests/test_gpu_methods.py -x -q > gpurun_out/c_pytest.log 2>&1
tail -5 gpurun_out/c_pytest.log
timeout 1200 python bench.py --verbose > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -c 1500 gpurun_out/c_bench.json; tail -3 gpurun_out/c_bench.err
