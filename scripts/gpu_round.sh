#!/bin/bash
mkdir -p gpurun_out
timeout 150 python scripts/ab_ranks.py 0 2 4 5 7 2>&1 | grep -v Warning | tee gpurun_out/r01d_ab_ranks.log
AURDF_ICP_SMALL=0 timeout 150 python scripts/ab_ranks.py 4 5 7 2>&1 | grep -v Warning | tee -a gpurun_out/r01d_ab_ranks.log
