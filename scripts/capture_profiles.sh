#!/bin/bash
# Captures the ncu evidence kept under profiles/ (run on the GPU box through gpurun):
#   1. launch list of the bench command (gpu__time_duration.sum per kernel, cold-cache, serialised)
#   2. one `--set full` capture of every kernel of one forward-shock grid step and one FS+RS series step
# usage: scripts/capture_profiles.sh TAG      -> gpurun_out/TAG_*.{csv,ncu-rep}
TAG=${1:-r01}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --inflight 1 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -s 22 -c 11 -o $O/${TAG}_full_fs \
    python scripts/prof_step.py fs 3 4096 > $O/${TAG}_full_fs.log 2>&1
ncu --set full --import-source on --clock-control none -s 24 -c 12 -o $O/${TAG}_full_rs \
    python scripts/prof_step.py rs 3 4096 > $O/${TAG}_full_rs.log 2>&1
tail -2 $O/${TAG}_full_fs.log $O/${TAG}_full_rs.log
