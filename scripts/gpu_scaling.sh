#!/bin/bash
# weak scaling of bench.py on one 8-GPU box (run under: gpurun --gpus 8)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
        bench.py --gpus $n --steps 50 --warmup 5 2>gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_$n.json"))
    print($n, "GPUs:", round(d["value"]), "frames/s  ms/step", round(d["ms_per_step"], 4), " e2e", round(d["e2e"]["value"]))
except Exception as e:
    print($n, "failed", e)
PY
done
