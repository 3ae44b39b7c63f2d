#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_pdl.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pdl.log
for pdl in 1 0; do
  CDR_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stream-microbench > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
  echo "bench pdl=$pdl rc=$?"
done
