#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_c.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest_c.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_c.err
