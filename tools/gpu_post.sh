#!/bin/bash
# GPU run: timing of fill_aux_buffers! / denoise! / postprocess! + the ncu launch list of their kernels (called through gpurun)
set -x
mkdir -p gpurun_out
timeout 300 python tools/post_bench.py 10 > gpurun_out/post_bench.json 2> gpurun_out/post_bench.err
cat gpurun_out/post_bench.json; tail -3 gpurun_out/post_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_aux|k_denoise|k_film_post|k_film_final" -c 60 --csv \
  --log-file gpurun_out/post_launches.csv python tools/post_bench.py 2 > gpurun_out/post_ncu.log 2>&1
tail -5 gpurun_out/post_ncu.log; wc -l gpurun_out/post_launches.csv
