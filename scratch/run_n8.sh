#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -c 600 gpurun_out/bench_n$N.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print('N=$N value',round(d['value']),'e2e',round(d['e2e']['value']),'bf16',round(d['bf16']['value']), 'bf16 e2e', round(d['bf16']['e2e']['value']))
print(json.dumps(d.get('full_pipeline'))[:400]); print(json.dumps(d.get('config5_full_pipeline_1024_pairs'))[:500])
P
