#!/bin/bash
cd fast-3d-human-pose-estimation_b200/csrc && ./build.sh -DCDR_ENABLE_PDL > /dev/null 2>&1; cd ../..
export CDR_PDL=1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_f.log
fails=0
for i in 1 2 3 4 5 6; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stream-microbench > gpurun_out/pdl_$i.json 2> gpurun_out/pdl_$i.err || fails=$((fails+1)); done
echo "PDL bench B=64 fails=$fails/6"
fails=0
for rep in 1 2 3 4 5 6; do timeout 200 python scratch/stress.py encoder 64 80 1000 > gpurun_out/st.log 2> gpurun_out/st.err || fails=$((fails+1)); done
echo "PDL encoder stress fails=$fails/6"
