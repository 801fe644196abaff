#!/bin/bash
# one gpurun call: GPU test suite, ncu launch list + full capture of the sweep kernels, bench (both arms),
# config sweep, per-call latency.   bash scripts/gpu_final.sh <tag>
TAG=${1:-r01c}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_small|box_count|mask_fill|tile_scan' -s 20 -c 5 -f \
    -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_prof.log 2>&1
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
timeout 900 python tests/measure/sweep_configs.py ${TAG} 2>&1 | tail -12
timeout 600 python tests/measure/call_latency.py ${TAG} 2>&1 | tail -5
cp profiles/${TAG}_configs.md profiles/${TAG}_call_latency.md gpurun_out/ 2>/dev/null
ls -la gpurun_out/
