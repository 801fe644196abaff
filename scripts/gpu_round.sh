#!/bin/bash
# one gpurun call: parity of the ICP path, then A/B timing of the small-tile kernel variants
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_icp_gpu.py -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_icp.log 2>&1
cat gpurun_out/pytest_icp.log
for v in 0 128 256; do
  AURDF_ICP_SMALL=$v timeout 300 python scripts/ab_small.py 2>&1 | grep -v Warning
done | tee gpurun_out/ab_small.log
AURDF_ICP_SMALL=128 timeout 300 python scripts/tile_latency.py 2>&1 | tail -28 | tee gpurun_out/tile_latency_128.log
AURDF_ICP_SMALL=256 timeout 300 python scripts/tile_latency.py 2>&1 | tail -12 | tee gpurun_out/tile_latency_256.log
