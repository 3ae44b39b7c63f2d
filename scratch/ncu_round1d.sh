#!/bin/bash
set -x
mkdir -p gpurun_out
COMMON="--steps 2 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline"
BF="python bench.py $COMMON --precision bf16"
FP="python bench.py $COMMON --precision fp32"
$BF > gpurun_out/plain_bf16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bf16.csv $BF > gpurun_out/ncu_launch_bf16.log 2>&1
$FP > gpurun_out/plain_fp32.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_fp32.csv $FP > gpurun_out/ncu_launch_fp32.log 2>&1
$BF > gpurun_out/plain_bf16b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_tc_kernel -s 4 -c 4 -o gpurun_out/prof_tc $BF > gpurun_out/ncu_tc.log 2>&1
$FP > gpurun_out/plain_fp32b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_tc_kernel -s 4 -c 4 -o gpurun_out/prof_f16x2 $FP > gpurun_out/ncu_f16x2.log 2>&1
ST="python scratch/stream_diag.py"
$ST > gpurun_out/plain_stream.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heat_stream -s 2 -c 1 -o gpurun_out/prof_heat $ST > gpurun_out/ncu_heat.log 2>&1
ENC="python scratch/enc_diag.py"
$ENC > gpurun_out/enc_diag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"tap_gemm_tc_kernel|stem_conv|maxpool" -s 0 -c 105 --csv --log-file gpurun_out/launches_encoder.csv $ENC > gpurun_out/ncu_launch_enc.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
