#!/bin/bash
# one gpurun call: GPU tests, config sweep, chamfer bench.  bash scripts/gpu_round.sh <tag>
TAG=${1:-r02d}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python tests/measure/sweep_configs.py ${TAG} 2>&1 | tail -16 | tee gpurun_out/${TAG}_sweep.log
cp profiles/${TAG}_configs.md gpurun_out/ 2>/dev/null
timeout 600 python scripts/bench_chamfer.py ${TAG} 2>&1 | tail -8 | tee gpurun_out/${TAG}_chamfer.log
cp profiles/${TAG}_chamfer.md gpurun_out/ 2>/dev/null
