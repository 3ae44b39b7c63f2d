#!/bin/bash
# CTA-pair (multicast A) variant of the f16x2 transposed convs: parity + A/B timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "decoder or stagewise or golden or full_size or f16x2" > gpurun_out/pytest_cl.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_cl.log
COMMON="--steps 20 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline"
for C in 1 0 1 0; do
  CDR_CLUSTER=$C timeout 300 python bench.py $COMMON > gpurun_out/bench_cl$C.json 2> gpurun_out/bench_cl$C.err; echo "bench CL=$C rc=$?"
  tail -c 300 gpurun_out/bench_cl$C.err
  python - <<P
import json
d=json.load(open('gpurun_out/bench_cl$C.json'))
print('CL=$C value',round(d['value']),'e2e',round(d['e2e']['value']),{k:round(v*1e3,1) for k,v in d['stages_ms'].items()})
P
done
