#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_sk.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sk.log
COMMON="--steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline --no-config5"
for SK in 1 0 1 0; do
  CDR_SPLITK=$SK timeout 300 python bench.py $COMMON > gpurun_out/bench_sk$SK.json 2> gpurun_out/bench_sk$SK.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_sk$SK.json'))
print('SK=$SK value',round(d['value']),'e2e',round(d['e2e']['value']),{k:round(v*1e3,1) for k,v in d['stages_ms'].items() if k in ('cf_conv1','deconv1','deconv2','deconv3')}, d['clocks'])
P
done
for cfg in "head_fp32 64 2000" "head_fp32 3 2000" "head_fp32 37 1000" "headgraph 64 1500"; do set -- $cfg; timeout 120 python scratch/stress.py $1 $2 $3 > gpurun_out/stsk.log 2> gpurun_out/stsk.err; echo "$cfg rc=$? $(tail -1 gpurun_out/stsk.log)"; done
