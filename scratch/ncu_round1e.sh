#!/bin/bash
# final round-1 evidence: GPU tests, default bench line, ncu launch lists, --set full captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
COMMON="--steps 2 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline"
BF="python bench.py $COMMON --precision bf16"
FP="python bench.py $COMMON --precision fp32"
$BF > gpurun_out/plain_bf16.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bf16.csv $BF > gpurun_out/ncu_launch_bf16.log 2>&1
$FP > gpurun_out/plain_fp32.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_fp32.csv $FP > gpurun_out/ncu_launch_fp32.log 2>&1
$BF > gpurun_out/plain_bf16b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_tc_kernel -s 4 -c 4 -o gpurun_out/prof_tc -f $BF > gpurun_out/ncu_tc.log 2>&1
$FP > gpurun_out/plain_fp32b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_tc_kernel -s 4 -c 4 -o gpurun_out/prof_f16x2 -f $FP > gpurun_out/ncu_f16x2.log 2>&1
ST="python scratch/stream_diag.py"
$ST > gpurun_out/plain_stream.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heat_stream -s 2 -c 1 -o gpurun_out/prof_heat -f $ST > gpurun_out/ncu_heat.log 2>&1
FT="python scratch/ftl_diag.py"
$FT > gpurun_out/plain_ftl.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ftl_vec -s 2 -c 2 -o gpurun_out/prof_ftl -f $FT > gpurun_out/ncu_ftl.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches*.csv
