mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_icp_gpu.py tests/test_integration_gpu.py -x -q -m gpu 2>&1 | tail -5 ) | tee gpurun_out/r02t_tests.log
timeout 600 python tests/measure/call_latency.py r02t 2>&1 | tail -5
AURDF_ICP_GRID=0 timeout 600 python tests/measure/call_latency.py r02t_nogrid 2>&1 | tail -5
AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee gpurun_out/r02t_bench.log
AURDF_ICP_SMALL_MINB=4 AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench minb4 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee -a gpurun_out/r02t_bench.log
AURDF_ICP_SMALL_MINB=6 AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench minb6 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])" | tee -a gpurun_out/r02t_bench.log
cp profiles/r02t*_call_latency.md gpurun_out/
