#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_final3.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref3.json 2>/dev/null; echo "ref rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_final3.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'bf16',round(d['bf16']['value']),'clocks',d['clocks'])
print('backward',json.dumps(d.get('backward_ops'))[:400])
print('c4',d['config4_softargmax_dlt_1m_poses']['poses_per_s'],'c5',d['config5_full_pipeline_1024_pairs']['pairs_per_s'])
r=json.load(open('gpurun_out/bench_ref3.json')); print('ref',r['value'],r['cpu_baseline']['cores'])
P
