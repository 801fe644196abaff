AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'])"
timeout 200 python scripts/tile_latency.py 2>&1 | grep -E "single tile max_iter=10000|all 900 tiles max_iter=10000|whole iteration|barrier A|totals"
