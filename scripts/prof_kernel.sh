#!/bin/bash
# One `--set full` capture (with source counters) of ONE kernel of the third step of a workload:
#   scripts/prof_kernel.sh TAG fs|rs BATCH KERNEL_REGEX  -> gpurun_out/TAG.ncu-rep
TAG=$1; W=$2; B=$3; K=$4
ncu --set full --import-source on --clock-control none --kernel-name regex:$K --launch-skip 2 --launch-count 1 \
    -o gpurun_out/$TAG python scripts/prof_step.py $W 3 $B > gpurun_out/$TAG.log 2>&1
tail -n 2 gpurun_out/$TAG.log
