#!/bin/bash
# ncu evidence for the bench command (called through gpurun); $1 = tag.
#  1. launch list of `python bench.py --steps 16 --warmup 16` (gpu__time_duration per launch)
#  2. --set full capture of the 12 k_trace launches of the timed pass (warm-up pass = launches 0..11)
#  3. --set full capture of the k_shade launches of the first 2 bounces of the timed pass
# The .ncu-rep files are summarised on the box (raw-page CSV + per-line hot spots) and only kept when small:
# gpurun copies back at most 64 MiB.
set -x
mkdir -p gpurun_out
T=${1:-x}
BENCH="python bench.py --quick --steps 16 --warmup 16"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $BENCH > gpurun_out/ncu_list_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 12 -c 12 -o /tmp/prof_trace_$T $BENCH > gpurun_out/ncu_trace_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 36 -c 6 -o /tmp/prof_shade_$T $BENCH > gpurun_out/ncu_shade_$T.log 2>&1
for K in trace shade; do
  R=/tmp/prof_${K}_$T.ncu-rep
  [ -f $R ] || continue
  ncu -i $R --page raw --csv > gpurun_out/${K}_raw_$T.csv 2>/dev/null
  python tools/ncu_summary.py $R > gpurun_out/${K}_summary_$T.txt 2>&1
  python tools/ncu_hot_lines.py $R 0 60 > gpurun_out/${K}_hot0_$T.txt 2>&1
  python tools/ncu_hot_lines.py $R 1 60 > gpurun_out/${K}_hot1_$T.txt 2>&1
  S=$(stat -c %s $R); if [ $S -lt 20000000 ]; then cp $R gpurun_out/; fi
done
ls -la gpurun_out /tmp/*.ncu-rep
