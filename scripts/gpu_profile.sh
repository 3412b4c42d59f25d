#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then ncu on the SAME command -- a launch list and full captures of
# the top kernels.  Usage (under gpurun): bash scripts/gpu_profile.sh [tag]
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32|tail_fused|splitk_reduce" -s 12 -c 16 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm profile rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:extract_tc -s 6 -c 2 -o gpurun_out/prof_extract -f $CMD > gpurun_out/ncu_extract.log 2>&1
echo "extract profile rc=$?"
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:gather_rows|clip_update|grad_finalize" -s 9 -c 4 -o gpurun_out/prof_elem -f $CMD > gpurun_out/ncu_elem.log 2>&1
echo "elementwise profile rc=$?"
ls -la gpurun_out | grep -E "ncu-rep|launches" | head
