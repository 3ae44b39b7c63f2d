#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "decoder or stagewise or golden or full_size or f16x2 or encoder or bf16" > gpurun_out/pytest_pf.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pf.log
COMMON="--steps 10 --warmup 3 --no-cpu-baseline --no-stream-microbench --no-full-pipeline --no-config5"
for PF in 1 0 1 0; do
  CDR_PREFETCH=$PF timeout 300 python bench.py $COMMON > gpurun_out/bench_pf$PF.json 2> gpurun_out/bench_pf$PF.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_pf$PF.json'))
f=lambda st:{k:round(v*1e3,1) for k,v in st.items() if 'deconv' in k or 'final' in k or 'cf_' in k}
print('PF=$PF fp32',round(d['value']),f(d['stages_ms']))
print('      bf16',round(d['bf16']['value']),f(d['bf16']['stages_ms']))
P
done
CDR_PREFETCH=1 timeout 200 python scratch/enc_diag.py 2>&1 | head -3
CDR_PREFETCH=0 timeout 200 python scratch/enc_diag.py 2>&1 | head -3
