#!/bin/bash
# A/B of library variants on the GPU box: scripts/ab.sh [fs|rs] BATCH KERNEL_REGEX lib1.so lib2.so ...
# prints k-kernel duration / instruction count (ncu, one launch) per variant built with different -D switches
W=$1; B=$2; K=$3; shift 3
for lib in "$@"; do
  echo "== $lib"
  VAG_LIB_PATH=$PWD/$lib ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_fp64.sum \
    --clock-control none --kernel-name regex:$K --launch-skip 2 --launch-count 1 python scripts/prof_step.py $W 3 $B 2>&1 | grep -E "gpu__time|inst_executed|wavefronts"
done
