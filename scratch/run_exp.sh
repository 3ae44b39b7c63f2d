#!/bin/bash
mkdir -p gpurun_out
COMMON="--steps 10 --warmup 3 --single-precision --no-cpu-baseline --no-stream-microbench --no-full-pipeline"
for V in base notmastore; do
  if [ $V = base ]; then unset CDR_LIB_PATH; else export CDR_LIB_PATH=$PWD/scratch/exp/lib_$V.so; fi
  timeout 300 python bench.py $COMMON > gpurun_out/bench_x$V.json 2> gpurun_out/bench_x$V.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_x$V.json'))
print('$V value',round(d['value']),{k:round(v*1e3,1) for k,v in d['stages_ms'].items() if 'deconv' in k or 'final' in k or 'conv1' in k}, 'mpjpe', d['mpjpe'])
P
done
