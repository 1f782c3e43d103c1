#!/bin/bash
# round 2, GPU call H: small-path tests, ncu of the tall predict kernel, trip kernel on CTA pairs (experiment)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small.py tests/test_gpu_ingest.py tests/test_gpu_cv.py -x -q > gpurun_out/h_pytest.log 2>&1
tail -5 gpurun_out/h_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 2 -c 1 -o gpurun_out/h_prof_predict python scripts/prof_predict.py > gpurun_out/h_ncu.log 2>&1
tail -2 gpurun_out/h_ncu.log
for tc in 0 1; do
  MBPLS_FUSED_TRIP_CLUSTER=$tc timeout 400 python scripts/bench_onepass.py 1.0 dense "v=one-pass trip+deflate" > gpurun_out/h_dense_tripcl${tc}.json 2>&1
  tail -1 gpurun_out/h_dense_tripcl${tc}.json | cut -c1-600
done
MBPLS_FUSED_TRIP_CLUSTER=1 timeout 600 python -m pytest tests/test_gpu_onepass.py -x -q > gpurun_out/h_pytest_tripcl.log 2>&1
tail -3 gpurun_out/h_pytest_tripcl.log
