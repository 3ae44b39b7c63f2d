#!/bin/bash
mkdir -p gpurun_out
COMMON="--steps 10 --warmup 3 --no-cpu-baseline --no-stream-microbench --no-full-pipeline --no-config5"
for V in base st12 base st12; do
  if [ $V = base ]; then unset CDR_LIB_PATH; else export CDR_LIB_PATH=$PWD/scratch/exp/lib_$V.so; fi
  timeout 300 python bench.py $COMMON > gpurun_out/bench_m$V.json 2> gpurun_out/bench_m$V.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_m$V.json'))
print('$V fp32',round(d['value']),'final',round(d['stages_ms']['final_1x1']*1e3,1),'| bf16',round(d['bf16']['value']),'final',round(d['bf16']['stages_ms']['final_1x1']*1e3,1), 'fusion', round(sum(v for k,v in d['bf16']['stages_ms'].items() if k.startswith('cf_'))*1e3,1))
P
done
for V in base st12; do
  if [ $V = base ]; then unset CDR_LIB_PATH; else export CDR_LIB_PATH=$PWD/scratch/exp/lib_$V.so; fi
  timeout 200 python scratch/enc_diag.py 2>&1 | head -2 | tail -1
done
unset CDR_LIB_PATH
timeout 120 python scratch/power_probe.py full_fp32 2>&1 | tail -3
