#!/bin/bash
# round-2 GPU run (through gpurun): $1 = tag; remaining args = what to do: tests | bench | prof:<config>:<kernel regex>:<skip>:<count>
set -x
mkdir -p gpurun_out
T=$1; shift
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for W in "$@"; do
  case $W in
    tests)
      timeout 1700 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/pytest_gpu_$T.full.log
      grep -E "within_tol|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_$T.full.log > gpurun_out/pytest_gpu_$T.log
      tail -120 gpurun_out/pytest_gpu_$T.full.log >> gpurun_out/pytest_gpu_$T.log; tail -5 gpurun_out/pytest_gpu_$T.log ;;
    bench)
      timeout 1200 python bench.py --steps 16 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 600 gpurun_out/bench_$T.json; tail -3 gpurun_out/bench_$T.err ;;
    benchref)
      timeout 1200 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; tail -c 600 gpurun_out/bench_ref_$T.json ;;
    cfg:*)
      IFS=: read -r _ CFG EXTRA <<< "$W"
      timeout 900 python bench.py --quick --no-per-config --config $CFG --steps 8 --warmup 3 $EXTRA > gpurun_out/cfg_${CFG}_$T.json 2> gpurun_out/cfg_${CFG}_$T.err; tail -c 400 gpurun_out/cfg_${CFG}_$T.json ;;
    list:*)
      IFS=: read -r _ CFG <<< "$W"
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${CFG}_$T.csv python bench.py --quick --no-per-config --config $CFG --steps 4 --warmup 3 > gpurun_out/ncu_list_${CFG}_$T.log 2>&1 ;;
    prof:*)
      IFS=: read -r _ CFG KRE SKIP CNT <<< "$W"
      R=/tmp/prof_${CFG}_${KRE//[^a-zA-Z0-9_]/}_$T
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o $R python bench.py --quick --no-per-config --config $CFG --steps 4 --warmup 3 > gpurun_out/ncu_${CFG}_$T.log 2>&1
      B=$(basename $R)
      python tools/ncu_summary.py $R.ncu-rep > gpurun_out/${B}_summary.txt 2>&1
      for i in $(seq 0 $((CNT-1))); do python tools/ncu_hot_lines.py $R.ncu-rep $i 50 > gpurun_out/${B}_hot$i.txt 2>&1; done
      S=$(stat -c %s $R.ncu-rep); if [ $S -lt 15000000 ]; then cp $R.ncu-rep gpurun_out/; fi ;;
    variants:*)
      IFS=: read -r _ CFGS STEPS <<< "$W"
      HK_BENCH_CONFIGS=$CFGS timeout 2400 python tools/variants.py run --steps ${STEPS:-8} > gpurun_out/variants_$T.log 2>&1; python tools/variants.py table | tee gpurun_out/variants_table_$T.txt ;;
    sanitize)
      timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_smoke_$T.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke_$T.out 2>&1
      timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_smoke_$T.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke_$T.out 2>&1
      tail -3 gpurun_out/sanitizer_*_$T.log ;;
  esac
done
ls -la gpurun_out | tail -20
