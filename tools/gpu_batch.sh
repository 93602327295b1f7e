#!/bin/bash
# $1 = tag; samples-in-flight sweep per config
mkdir -p gpurun_out
T=$1; shift
for spec in "$@"; do
  IFS=: read -r CFG B <<< "$spec"
  timeout 900 python bench.py --quick --no-per-config --config $CFG --steps 16 --warmup 3 --batch $B > gpurun_out/batch_${CFG}_${B}_$T.json 2> gpurun_out/batch_${CFG}_${B}_$T.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/batch_${CFG}_${B}_$T.json").read().strip().splitlines()[-1])
    print("$CFG batch $B: %.1f Msamples/s %.3f ms/step e2e %s" % (j["value"], j["ms_per_step"], j["e2e"] and round(j["e2e"]["value"], 1)), {k: round(v, 2) for k, v in j["roofline"]["stage_ms_per_step"].items()})
except Exception as e:
    print("$CFG batch $B failed", e, open("gpurun_out/batch_${CFG}_${B}_$T.err").read()[-400:])
PY
done
