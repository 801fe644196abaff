#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/pytest_gpu.log
for mb in 5 6 7; do
  AURDF_ICP_SMALL_MINB=$mb timeout 300 python scripts/ab_small.py 2>&1 | grep -v Warning
done | tee gpurun_out/ab_small.log
AURDF_ICP_SMALL=0 timeout 300 python scripts/ab_small.py 2>&1 | grep -v Warning | tee -a gpurun_out/ab_small.log
for ch in 1 3; do
  echo "== AURDF_HOST_CHUNKS=$ch"
  AURDF_HOST_CHUNKS=$ch timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1
done | tee gpurun_out/bench_chunks.log
