#!/bin/bash
# Round-2 profiling pass of the f16x2 (reference-precision) encoder — run under gpurun, ONE GPU.  Every ncu command follows
# the same command run plain (exit 0).  Raw pages are summarised by profiles/summarize_r02_encoder.py.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
S="python devtools/enc_diag.py fp32 128"
$S > $O/plain_enc_fp32.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:tap_gemm|stem_conv|maxpool' -c 700 --csv --log-file $O/launches_enc_fp32.csv $S > $O/ncu_enc_launches.log 2>&1
# one forward = 2 + 103 kernels; enc_diag runs 3 warm-ups first: skip them, capture stem + pool + layer1 + layer2 + the
# first blocks of layer3 (every kernel shape of the encoder occurs in that window)
$S > /dev/null 2>&1 &&
ncu --set full --clock-control none -k 'regex:tap_gemm|stem_conv|maxpool' -s 315 -c 36 -f -o /tmp/prof_enc_fp32 $S > $O/ncu_enc_full.log 2>&1
ncu -i /tmp/prof_enc_fp32.ncu-rep --page raw --csv > $O/prof_enc_fp32.raw.csv 2>/dev/null
du -sh $O
