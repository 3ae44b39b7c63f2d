#!/bin/bash
mkdir -p gpurun_out
COMMON="--steps 10 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline --precision bf16"
for BN in 256 128 256 128; do
  CDR_BF16_DECONV_BN=$BN timeout 300 python bench.py $COMMON > gpurun_out/bench_bn$BN.json 2> gpurun_out/bench_bn$BN.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_bn$BN.json'))
print('BN=$BN value',round(d['value']),{k:round(v*1e3,1) for k,v in d['stages_ms'].items() if 'deconv' in k or 'final' in k})
P
done
