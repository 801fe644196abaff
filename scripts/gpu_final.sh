#!/bin/bash
# one gpurun call: GPU test suite, ncu launch list + full captures (sweep kernels on C2, the grid kernel's cluster
# variant on a C5 sweep point, the fused chamfer kernel), bench (both arms), config sweep, per-call latency,
# chamfer bench.   bash scripts/gpu_final.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'icp_small|box_count|mask_fill|tile_scan|icp_grid' -s 24 -c 6 -f \
    -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'icp_grid_kernel|grid_build' -s 4 -c 2 -f \
    -o gpurun_out/${TAG}_prof_c5 python scripts/profile_c5.py 65536 8 4 > gpurun_out/${TAG}_prof_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'chamfer_fwd|chamfer_bwd' -s 8 -c 2 -f \
    -o gpurun_out/${TAG}_prof_chamfer python scripts/profile_chamfer.py > gpurun_out/${TAG}_prof_chamfer.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
timeout 900 python tests/measure/sweep_configs.py ${TAG} 2>&1 | tail -14
timeout 600 python tests/measure/call_latency.py ${TAG} 2>&1 | tail -5
timeout 600 python scripts/bench_chamfer.py ${TAG} 2>&1 | tail -6
cp profiles/${TAG}_configs.md profiles/${TAG}_call_latency.md profiles/${TAG}_chamfer.md gpurun_out/ 2>/dev/null
ls -la gpurun_out/ | tail -30
