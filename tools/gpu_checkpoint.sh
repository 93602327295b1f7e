#!/bin/bash
# full GPU suite + smoke + the two bench arms (what the driver runs at round end)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
( time python bench.py ) > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 400 gpurun_out/r02_bench_c3.json; tail -4 gpurun_out/r02_bench_c3.err
( time python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 300 gpurun_out/r02_bench_reference.json
