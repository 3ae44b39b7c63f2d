#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_enc.log 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|tcgen05|full pipeline|bf16 head|Error|stage taps" gpurun_out/pytest_enc.log | head -20
timeout 300 python scratch/enc_diag.py > gpurun_out/enc_diag.log 2>&1
echo "diag rc=$?"
cat gpurun_out/enc_diag.log | head -60
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "bench rc=$?"
