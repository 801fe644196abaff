#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_icp_gpu.py -x -q 2>&1 | tail -6 ) | tee gpurun_out/r01d_pytest_icp.log
timeout 200 python scripts/ab_small.py 2>&1 | grep -v Warning | tee gpurun_out/r01d_ab_split1.log
AURDF_ICP_SMALL_SPLIT=0 timeout 200 python scripts/ab_small.py franka 2>&1 | grep -v Warning | tee gpurun_out/r01d_ab_split0.log
timeout 300 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/r01d_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r01d_bench.json')); print('bench', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])"
