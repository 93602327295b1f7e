#!/bin/bash
# ncu evidence for the bench command (called through gpurun); $1 = tag.
#  1. launch list of `python bench.py --steps 16 --warmup 16` (gpu__time_duration per launch)
#  2. --set full capture of the 12 k_trace launches of the timed pass (warm-up pass = launches 0..11)
#  3. --set full capture of the k_shade launches of the first 4 bounces of the timed pass
set -x
mkdir -p gpurun_out
T=${1:-x}
BENCH="python bench.py --quick --steps 16 --warmup 16"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $BENCH > gpurun_out/ncu_list_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 12 -c 12 -o gpurun_out/prof_trace_$T $BENCH > gpurun_out/ncu_trace_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 36 -c 12 -o gpurun_out/prof_shade_$T $BENCH > gpurun_out/ncu_shade_$T.log 2>&1
ls -la gpurun_out
