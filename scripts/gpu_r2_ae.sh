#!/bin/bash
# round 2, GPU call AE (last): ncu --set full of the final non-headline kernels: the TMA-ring tall product (predict at the C5 shape) and the
# FP64 tensor-core cross product (X'X at the C5 shape)
mkdir -p gpurun_out
timeout 110 ncu --set full --clock-control none --import-source on -k regex:"skinny_tall" -s 1 -c 1 -f -o gpurun_out/ae_prof_predict python scripts/prof_predict.py > gpurun_out/ae_ncu_predict.log 2>&1
tail -2 gpurun_out/ae_ncu_predict.log
timeout 130 ncu --set full --clock-control none --import-source on -k regex:"crossprod" -s 1 -c 1 -f -o gpurun_out/ae_prof_crossprod python scripts/bench_crossprod.py > gpurun_out/ae_ncu_crossprod.log 2>&1
tail -3 gpurun_out/ae_ncu_crossprod.log
ls -la gpurun_out/ae_*.ncu-rep
