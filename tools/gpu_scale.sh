#!/bin/bash
# multi-GPU bench lines (called through gpurun --gpus 8): $@ = CFG:N specs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt
P=29500
for spec in "$@"; do
  IFS=: read -r CFG N <<< "$spec"
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 16 --warmup 3 --config $CFG > gpurun_out/scale_${CFG}_n$N.json 2> gpurun_out/scale_${CFG}_n$N.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/scale_${CFG}_n$N.json").read().strip().splitlines()[-1])
    print("$CFG N=$N: %.1f Msamples/s %.3f ms/step reduce %.2f ms e2e %.1f" % (j["value"], j["ms_per_step"], j["film_reduce_ms"], j["e2e"]["value"]))
except Exception as e:
    print("$CFG N=$N failed", e, open("gpurun_out/scale_${CFG}_n$N.err").read()[-600:])
PY
done
