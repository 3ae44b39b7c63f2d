#!/bin/bash
mkdir -p gpurun_out
for M in 1 2; do
CDR_LSU_STORE=$M timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_lsu$M.log 2>&1; echo "pytest LSU=$M rc=$?"; tail -2 gpurun_out/pytest_lsu$M.log
done
COMMON="--steps 10 --warmup 3 --no-cpu-baseline --no-stream-microbench --no-full-pipeline --no-config5"
for M in 0 1 2 0 1 2; do
  CDR_LSU_STORE=$M timeout 300 python bench.py $COMMON > gpurun_out/bench_lsu$M.json 2> gpurun_out/bench_lsu$M.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_lsu$M.json'))
f=lambda st:{k:round(v*1e3,1) for k,v in st.items() if 'deconv' in k or 'final' in k or 'cf_' in k}
print('LSU=$M fp32',round(d['value']),f(d['stages_ms']))
print('       bf16',round(d['bf16']['value']),f(d['bf16']['stages_ms']))
P
done
for M in 0 2; do CDR_LSU_STORE=$M timeout 200 python scratch/enc_diag.py 2>&1 | head -2 | tail -1; done
