#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14
COMMON="--gpus 8 --steps 20 --warmup 3 --no-full-pipeline --single-precision --no-cpu-baseline"
for NB in 0 1 0 1; do
  CDR_NO_NUMA_BIND=$NB timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$NB bench.py $COMMON > gpurun_out/bench_numa$NB.json 2> gpurun_out/bench_numa$NB.err; echo "rc=$?"
  python - <<P
import json
d=json.loads([l for l in open('gpurun_out/bench_numa$NB.json') if l.startswith('{')][-1])
print('NO_BIND=$NB value',round(d['value']),'e2e',round(d['e2e']['value']), d['config'].get('host_binding'))
P
done
