#!/bin/bash
# one gpurun call for the grid-search kernel: GPU tests, config sweep, ncu of one C5 point.  bash scripts/gpu_grid.sh <tag>
TAG=${1:-r02b}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python tests/measure/sweep_configs.py ${TAG} 2>&1 | tail -16 | tee gpurun_out/${TAG}_sweep.log
cp profiles/${TAG}_configs.md gpurun_out/ 2>/dev/null
