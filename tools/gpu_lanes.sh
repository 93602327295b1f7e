#!/bin/bash
# four-lane frame pipelining: parity tests for the lanes + the e2e loop with 1..3 read-outs left in flight
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "pipelin or stream or wave or error or aux" > gpurun_out/lanes_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/lanes_tests.log
for D in 3 2 1; do
  echo "== e2e depth $D"
  HK_E2E_DEPTH=$D timeout 300 python tools/diag_e2e.py C3 2>&1 | grep -E " (B|D) "
done
echo "== connections 8, depth 3"; CUDA_DEVICE_MAX_CONNECTIONS=8 HK_E2E_DEPTH=3 timeout 300 python tools/diag_e2e.py C3 2>&1 | grep -E " (B|D) "
echo "== connections 64, depth 3"; CUDA_DEVICE_MAX_CONNECTIONS=64 HK_E2E_DEPTH=3 timeout 300 python tools/diag_e2e.py C3 2>&1 | grep -E " (B|D) "
for c in C2 C4 C5; do HK_E2E_DEPTH=3 timeout 300 python tools/diag_e2e.py $c 2>&1 | grep -E " (A|A16|B|D) "; done
