#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_coord_map.py -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/pytest_coord_map.log
timeout 300 python scripts/bench_coord_map.py r01 2>&1 | tail -6
cp profiles/r01_coord_map.md gpurun_out/
