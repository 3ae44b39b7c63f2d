#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_g.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"
tail -c 800 gpurun_out/bench_n2b.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2b.json') if l.startswith('{')][-1])
print('N=2 value',round(d['value']),'e2e',round(d['e2e']['value']),'bf16',round(d['bf16']['value']))
print(json.dumps(d.get('full_pipeline'))[:600]); print(json.dumps(d.get('config5_full_pipeline_1024_pairs'))[:900])
P
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | tail -c 700
