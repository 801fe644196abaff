#!/bin/bash
# A/B of the host path's frame blocks (one stream each) on wx200_5.   usage: scripts/ab_chunks.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
for rep in 1 2; do
for ch in 1 2 3 4; do
  AURDF_BENCH_SKIP_CPU=1 AURDF_HOST_CHUNKS=$ch timeout 120 python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/chunks_$ch.json
  python -c "
import json; d=json.load(open('gpurun_out/chunks_$ch.json')); print('chunks $ch: e2e', round(d['e2e']['value']), ' value', round(d['value']))"
done; done | tee gpurun_out/${TAG}_chunks.log
