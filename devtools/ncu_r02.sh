#!/bin/bash
# Round-2 profiling pass (run under gpurun, ONE GPU): every ncu command follows the same command run plain (exit 0).
# The .ncu-rep files are summarised ON THE BOX (raw-page csv) and removed: gpurun_out/ may carry 64 MiB back.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
K='regex:pinv2|amax|nchw_to_rows|tap_gemm|ftl|deconv_tail|tail_merge|heat_stream'
for p in fp32 bf16; do
  B="python bench.py --steps 2 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline --no-sustained --precision $p"
  $B > $O/plain_bench_$p.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$p.csv $B > $O/ncu_bench_$p.log 2>&1
  S="python devtools/stage_times.py $p 64 2"
  $S > $O/plain_stage_$p.log 2>&1 &&
  # 3 warm-up steps of 12 kernels of this library (both modes: the fp32 layout pass is one launch now): start at a step boundary
  SKIP=36
  ncu --set full --clock-control none -k "$K" -s $SKIP -c 12 -f -o /tmp/prof_step_$p $S > $O/ncu_step_$p.log 2>&1
  ncu -i /tmp/prof_step_$p.ncu-rep --page raw --csv > $O/prof_step_$p.raw.csv 2>/dev/null
  $S > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:deconv_tail -s 3 -c 1 -f -o $O/prof_tail_$p $S > $O/ncu_tail_$p.log 2>&1
  ncu -i $O/prof_tail_$p.ncu-rep --page source --csv > $O/prof_tail_$p.source.csv 2>/dev/null
done
rm -f $O/*.log.tmp
du -sh $O
