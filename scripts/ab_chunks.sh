#!/bin/bash
mkdir -p gpurun_out
for ch in 4 2 3; do
  AURDF_BENCH_SKIP_CPU=1 AURDF_HOST_CHUNKS=$ch timeout 120 python bench.py --steps 80 --warmup 5 2>/dev/null | tail -1 > gpurun_out/chunks_$ch.json
  python -c "
import json; d=json.load(open('gpurun_out/chunks_$ch.json')); print('chunks $ch: e2e', round(d['e2e']['value']), ' value', round(d['value']))"
done | tee gpurun_out/r01e_chunks.log
