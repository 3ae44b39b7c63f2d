#!/bin/bash
run() {
  local tag="$1"; shift
  local fails=0
  for rep in 1 2 3 4 5 6 7 8 9 10; do
    env "$@" timeout 200 python scratch/stress.py encoder 64 80 1000 > gpurun_out/st.log 2> gpurun_out/st.err || fails=$((fails+1))
  done
  echo "$tag fails=$fails/10"
}
run "conv3+ds" CDR_ENC_SKIP=3
run "all" A=1
