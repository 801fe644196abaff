#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) | tee gpurun_out/r01e_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r01e_smoke.log
timeout 300 python scripts/call_latency.py r01e 2>&1 | tail -5
cp profiles/r01e_call_latency.md gpurun_out/ 2>/dev/null
