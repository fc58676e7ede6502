#!/bin/bash
# Round-2 profiling recipe (runs under gpurun on one B200): plain runs first, then the ncu
# launch list (warm caches) and `ncu --set full` captures of the hot kernels
# (B200_PROFILING.md).  Outputs land in gpurun_out/ and are summarised into profiles/r02/ by
# profiles/summarize_r02.py.
set -x
mkdir -p gpurun_out
export CDR_NO_CUDA_GRAPH=1
CMD="python bench.py --no-stress --no-kmeans --steps 3 --warmup 3 --cpu-steps 0 --min-timed-ms 1"
$CMD > gpurun_out/r02p_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 4000 --csv \
    --log-file gpurun_out/r02p_launches.csv $CMD > gpurun_out/r02p_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'reduce_samples_tma_kernel|reduce_features_strip_kernel|gpnh_weights_fused_kernel|aa_weights_fused_kernel|aa_head_kernel|aa_finalize_ls_kernel|aa_kzt_gradient_kernel' \
    -s 60 -c 9 -o gpurun_out/r02p_aa $CMD --workload aa > gpurun_out/r02p_ncu_aa.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'reduce_samples_tma_kernel|reduce_features_strip_kernel|gpnh_weights_fused_kernel' \
    -s 30 -c 3 -o gpurun_out/r02p_gpnh $CMD --workload gpnh > gpurun_out/r02p_ncu_gpnh.log 2>&1
# SKIP_TENSOR=1: only the iteration kernels (the SYRK / k = 64 kernels did not change)
[ "$SKIP_TENSOR" = "1" ] && { ls -la gpurun_out | grep r02p; exit 0; }
python profiles/bench_gram.py > gpurun_out/r02p_gram_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'syrk_tile_kernel' -s 2 -c 1 \
    -o gpurun_out/r02p_syrk python profiles/bench_gram.py > gpurun_out/r02p_ncu_syrk.log 2>&1
python profiles/bench_stress_kernels.py 18000 64 > gpurun_out/r02p_stress_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'features64_kernel|samples64_kernel' \
    -s 3 -c 3 -o gpurun_out/r02p_gemm64 python profiles/bench_stress_kernels.py 18000 64 \
    > gpurun_out/r02p_ncu_gemm64.log 2>&1
ls -la gpurun_out | grep r02p
