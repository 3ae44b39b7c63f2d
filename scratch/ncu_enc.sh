#!/bin/bash
mkdir -p gpurun_out
python scratch/enc_diag.py > gpurun_out/enc_diag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tap_gemm_tc_kernel -s 4 -c 26 -o gpurun_out/prof_enc python scratch/enc_diag.py > gpurun_out/ncu_enc.log 2>&1
ls -la gpurun_out/prof_enc.ncu-rep
