#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_c.log 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|stage taps|CUDA fp32 vs|full pipeline \[|Error" gpurun_out/pytest_c.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "bench rc=$?"
tail -c 300 gpurun_out/bench_c.err
