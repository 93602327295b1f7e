#!/bin/bash
# $1 = tag
set -x
mkdir -p gpurun_out
T=$1
timeout 900 python tools/diag_c4_batch.py > gpurun_out/diag_c4_$T.log 2>&1
grep -v "^    " gpurun_out/diag_c4_$T.log | tail -40
