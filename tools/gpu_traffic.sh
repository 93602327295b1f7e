#!/bin/bash
# ncu --set full capture of the k_trace launches of the timed pass of the bench command, per config ($@), + the DRAM traffic JSON
mkdir -p gpurun_out
for CFG in "$@"; do
  D=12; [ $CFG == C4 ] && D=32; [ $CFG == C5 ] && D=8; [ $CFG == C1 ] && D=5
  CMD="python bench.py --quick --no-per-config --config $CFG --steps 16 --warmup 3"
  R=/tmp/prof_trace_$CFG
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^k_trace -s $D -c $D -o $R $CMD > gpurun_out/ncu_trace_$CFG.log 2>&1
  python tools/ncu_summary.py $R.ncu-rep > gpurun_out/r02_${CFG}_k_trace_ncu_summary.txt 2>&1
  python tools/ncu_hot_lines.py $R.ncu-rep 0 40 > gpurun_out/r02_${CFG}_k_trace_hot_lines_bounce0.txt 2>&1
  python tools/ncu_hot_lines.py $R.ncu-rep 1 40 > gpurun_out/r02_${CFG}_k_trace_hot_lines_bounce1.txt 2>&1
  python tools/trace_traffic.py $R.ncu-rep $CFG "ncu --set full --clock-control none -k regex:^k_trace -s $D -c $D $CMD (the $D k_trace launches of the timed pass, 16 samples in flight)"
done
