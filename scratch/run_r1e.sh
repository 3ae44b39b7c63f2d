#!/bin/bash
# round-1 session 3: heat-kernel warp sweep, vector FTL, config 4 / config 5 bench fields
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_e.log
for W in 8 10 12; do
  echo "== CDR_HEAT_WARPS=$W"; CDR_HEAT_WARPS=$W timeout 300 python scratch/stream_diag.py 2>&1 | tail -6
done > gpurun_out/heat_sweep.log 2>&1
cat gpurun_out/heat_sweep.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_e.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_e.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'bf16',d['bf16']['value'])
print('stages',d['stages_ms'])
for k in ('roofline_hbm_stream','config4_softargmax_dlt_1m_poses','roofline_hbm_ftl','config5_full_pipeline_1024_pairs'):
    print(k, json.dumps(d.get(k))[:900])
P
