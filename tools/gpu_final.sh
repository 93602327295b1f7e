#!/bin/bash
mkdir -p gpurun_out
time python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 300 gpurun_out/r02_bench_c3.json; tail -2 gpurun_out/r02_bench_c3.err
time python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 300 gpurun_out/r02_bench_reference.json
R=/tmp/prof_lbvh
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_hit_lights_bvh -s 12 -c 2 -o $R python bench.py --quick --no-per-config --config C3 --steps 16 --warmup 3 > gpurun_out/ncu_lbvh.log 2>&1
python tools/ncu_summary.py $R.ncu-rep > gpurun_out/r02_c3_k_hit_lights_bvh_ncu_summary.txt 2>&1
python tools/ncu_hot_lines.py $R.ncu-rep 0 30 > gpurun_out/r02_c3_k_hit_lights_bvh_hot_lines.txt 2>&1
grep -E "^time|active lanes|issue active|occupancy|top stalls" gpurun_out/r02_c3_k_hit_lights_bvh_ncu_summary.txt | head -6
