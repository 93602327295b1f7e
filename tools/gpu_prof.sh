#!/bin/bash
# ncu captures of the dominant kernels (called through gpurun); $1 = tag
set -x
mkdir -p gpurun_out
T=${1:-x}
# bounce 0, 1, 3 of the 4th sample pass: launches are per pass 12 x k_trace; skip 3 passes (warmup) = 36
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 36 -c 4 -o gpurun_out/prof_trace_$T python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_trace_$T.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 108 -c 6 -o gpurun_out/prof_shade_$T python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_shade_$T.log 2>&1
ls -la gpurun_out
