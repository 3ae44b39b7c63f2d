#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -k "encoder or pipeline or poseresnet or frames" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_f.log
fails=0
for rep in 1 2 3 4 5 6 7 8; do timeout 200 python scratch/stress.py encoder 64 80 1000 > gpurun_out/st.log 2> gpurun_out/st.err || fails=$((fails+1)); done
echo "encoder stress fails=$fails/8"
timeout 300 python scratch/enc_diag.py > gpurun_out/enc_diag.log 2>&1; head -3 gpurun_out/enc_diag.log; grep -E "block(1|4|8|31)\.conv3|block8\.conv[12]" gpurun_out/enc_diag.log
