mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_integration_gpu.py tests/test_icp_gpu.py -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/r02f_tests.log
AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee gpurun_out/r02f_bench.log
for wl in franka allegro_hand; do AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 30 --warmup 5 --workload $wl 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee -a gpurun_out/r02f_bench.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_grid_kernel -s 3 -c 1 -f -o gpurun_out/r02f_grid python scripts/profile_c5.py 16384 32 > gpurun_out/r02f_grid_prof.log 2>&1
tail -3 gpurun_out/r02f_grid_prof.log
