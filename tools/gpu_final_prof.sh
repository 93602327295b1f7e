#!/bin/bash
# round-2 evidence for profiles/: launch lists of the bench command per config + ncu --set full of the largest kernels (current code)
mkdir -p gpurun_out
for CFG in C3 C4 C5 C2; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_${CFG}.csv python bench.py --quick --no-per-config --config $CFG --steps 16 --warmup 3 > gpurun_out/ncu_list_${CFG}.log 2>&1
  python tools/launch_summary.py gpurun_out/r02_launches_${CFG}.csv > gpurun_out/r02_launches_${CFG}_summary.txt 2>&1
  echo == $CFG; head -12 gpurun_out/r02_launches_${CFG}_summary.txt
done
prof() {  # cfg kernel-regex skip count tag
  R=/tmp/prof_$5
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $R python bench.py --quick --no-per-config --config $1 --steps 16 --warmup 3 > gpurun_out/ncu_$5.log 2>&1
  python tools/ncu_summary.py $R.ncu-rep > gpurun_out/r02_$5_ncu_summary.txt 2>&1
  python tools/ncu_hot_lines.py $R.ncu-rep 0 40 > gpurun_out/r02_$5_hot_lines.txt 2>&1
  grep -E "^kernel  |^time|active lanes|issue active|occupancy|top stalls" gpurun_out/r02_$5_ncu_summary.txt | head -14
}
prof C3 k_hit_lights 12 2 c3_k_hit_lights
prof C3 "k_shade" 36 3 c3_k_shade
prof C5 "k_shade<\(int\)5" 16 2 c5_k_shade_coated
prof C4 k_medium_track 33 2 c4_k_medium_track
prof C4 k_shadow_seg_ratio 320 2 c4_k_shadow_seg_ratio
