#!/bin/bash
# ncu --set full of the layered shading kernel on C5 (coated diffuse, part 1) and the matte kernel on C2
mkdir -p gpurun_out
prof() {  # cfg kernel-regex skip count tag
  R=/tmp/prof_$5
  timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$2" -s $3 -c $4 -o $R python bench.py --quick --no-per-config --config $1 --steps 16 --warmup 3 > gpurun_out/ncu_$5.log 2>&1
  python tools/ncu_summary.py $R.ncu-rep > gpurun_out/r02_$5_ncu_summary.txt 2>&1
  python tools/ncu_hot_lines.py $R.ncu-rep 0 45 > gpurun_out/r02_$5_hot_lines.txt 2>&1
  grep -E "^kernel  |^time|active lanes|issue active|occupancy|top stalls|^regs" gpurun_out/r02_$5_ncu_summary.txt | head -8
}
prof C5 'k_shade<.int.5, .bool.0, .int.1' 1 1 c5_k_shade_coated
prof C2 'k_shade<.int.1,' 1 1 c2_k_shade_matte
