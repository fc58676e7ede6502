#!/bin/bash
# Runs under gpurun: plain bench first, then the ncu launch list and one full capture of
# the streaming kernels and the batched QP kernel (B200_PROFILING.md recipe).
set -x
mkdir -p gpurun_out
export CDR_NO_CUDA_GRAPH=1
CMD="python bench.py --steps 2 --warmup 1 --cpu-steps 0"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'reduce_samples_tma_kernel|reduce_features_strip_kernel|qp_batched_kernel' -s 3 -c 6 \
    -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
