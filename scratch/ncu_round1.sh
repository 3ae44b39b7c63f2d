#!/bin/bash
# ncu evidence for round 1 (run under gpurun). Each ncu run follows a plain run of the same command.
set -x
mkdir -p gpurun_out
BF="python bench.py --steps 2 --warmup 3 --precision bf16 --no-cpu-baseline"
FP="python bench.py --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline"
$BF > gpurun_out/plain_bf16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bf16.csv $BF > gpurun_out/ncu_launch_bf16.log 2>&1
$BF > gpurun_out/plain_bf16b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_tc_kernelILi256 -s 3 -c 4 -o gpurun_out/prof_tc256 $BF > gpurun_out/ncu_tc.log 2>&1
$BF > gpurun_out/plain_bf16c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heat_stream -c 1 -o gpurun_out/prof_heat $BF > gpurun_out/ncu_heat.log 2>&1
$FP > gpurun_out/plain_fp32.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fp32.csv $FP > gpurun_out/ncu_launch_fp32.log 2>&1
$FP > gpurun_out/plain_fp32b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_ffma_kernelILi128 -s 5 -c 1 -o gpurun_out/prof_ffma $FP > gpurun_out/ncu_ffma.log 2>&1
ls -la gpurun_out/
