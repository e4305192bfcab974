#!/bin/bash
# Captures the ncu evidence kept under profiles/ (run on the GPU box through gpurun):
#   1. launch list of the bench command (gpu__time_duration.sum per kernel, cold-cache, serialised)
#   2. one `--set full` capture of every kernel of one forward-shock grid step (bench batch) and of one
#      FS+RS series step (config-5 batch); 10 kernel launches per step, the third step is captured
# usage: scripts/capture_profiles.sh TAG [FS_BATCH] [RS_BATCH]   -> gpurun_out/TAG_*.{csv,ncu-rep}
TAG=${1:-r01}
FSB=${2:-16384}
RSB=${3:-4096}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --inflight 1 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --launch-skip 20 --launch-count 10 -o $O/${TAG}_full_fs \
    python scripts/prof_step.py fs 3 $FSB > $O/${TAG}_full_fs.log 2>&1
ncu --set full --clock-control none --launch-skip 20 --launch-count 10 -o $O/${TAG}_full_rs \
    python scripts/prof_step.py rs 3 $RSB > $O/${TAG}_full_rs.log 2>&1
tail -n 2 $O/${TAG}_full_fs.log $O/${TAG}_full_rs.log
# summaries are produced on the box; the raw reports are dropped when they would not fit the 64 MiB
# return channel of gpurun
for f in fs rs; do
  python scripts/ncu_summary.py $O/${TAG}_full_$f.ncu-rep $O/${TAG}_full_${f}_summary.csv
  python scripts/sass_hist.py $O/${TAG}_full_$f.ncu-rep k_eats 24 > $O/${TAG}_sass_hist_eats_$f.txt
  python scripts/sass_hist.py $O/${TAG}_full_$f.ncu-rep k_dynamics 24 > $O/${TAG}_sass_hist_dynamics_$f.txt
done
python scripts/sass_hist.py $O/${TAG}_full_fs.ncu-rep k_grid 24 > $O/${TAG}_sass_hist_grid_fs.txt
du -sm $O/*.ncu-rep
for r in $O/${TAG}_full_fs.ncu-rep $O/${TAG}_full_rs.ncu-rep; do
  if [ $(du -sm $O | cut -f1) -gt 55 ]; then rm -f $r; fi
done
