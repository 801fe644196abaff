mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_icp_gpu.py -x -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/r02s_tests.log
for c in "16384 32" "65536 128"; do timeout 300 python scripts/grid_stats.py $c 2>&1 | grep -E "^C5|longest|sweep" | tee -a gpurun_out/r02s_grid_stats.log; done
AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee gpurun_out/r02s_bench.log
AURDF_ICP_GRID=0 AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench (no grid kernel) value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee -a gpurun_out/r02s_bench.log
