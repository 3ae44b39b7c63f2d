#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "encoder or full_pipeline" > gpurun_out/pytest_enc.log 2>&1
echo "pytest rc=$?"
tail -30 gpurun_out/pytest_enc.log
