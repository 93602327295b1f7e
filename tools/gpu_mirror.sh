#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -k "media or nanovdb or cloud or c4 or C4 or majorant or update_medium or fog or nebula or smoke or wave" 2>&1 | tail -4
echo "== mirror on"; bash tools/gpu_batch.sh mirror_on C4:16
echo "== mirror off"; HK_NO_DENSE_MIRROR=1 bash tools/gpu_batch.sh mirror_off C4:16
