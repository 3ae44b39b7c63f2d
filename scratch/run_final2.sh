#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final_$i.log 2>&1; echo "pytest run $i rc=$?"; tail -1 gpurun_out/pytest_final_$i.log; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"
BW="python scratch/bwd_diag.py"
$BW > gpurun_out/plain_bwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:softargmax_backward -s 2 -c 1 -o gpurun_out/prof_bwd -f $BW > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/prof_bwd.ncu-rep
timeout 600 python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_final4.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'bf16',round(d['bf16']['value']),'clocks',d['clocks'])
P
