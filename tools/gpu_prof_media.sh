#!/bin/bash
# ncu --set full of the two media kernels on C4 (bounce-1 launches of the first pass) and of k_camera / k_escaped on C3
mkdir -p gpurun_out
prof() {  # cfg kernel-regex skip count tag
  R=/tmp/prof_$5
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $R python bench.py --quick --no-per-config --config $1 --steps 16 --warmup 3 > gpurun_out/ncu_$5.log 2>&1
  python tools/ncu_summary.py $R.ncu-rep > gpurun_out/r02_$5_ncu_summary.txt 2>&1
  python tools/ncu_hot_lines.py $R.ncu-rep 0 45 > gpurun_out/r02_$5_hot_lines.txt 2>&1
  grep -E "^kernel  |^time|active lanes|issue active|occupancy|top stalls|^regs" gpurun_out/r02_$5_ncu_summary.txt | head -8
}
prof C4 k_medium_track 1 1 c4b_k_medium_track
prof C4 k_shadow_seg_ratio 10 1 c4b_k_shadow_seg_ratio
prof C3 k_camera 1 1 c3_k_camera
prof C3 k_escaped 13 1 c3_k_escaped
