mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) | tee gpurun_out/r02u_tests.log
AURDF_BENCH_SKIP_CPU=1 timeout 200 python bench.py --steps 50 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', d['roofline']['kernel_ms'], 'ms_per_step', d['ms_per_step'])" | tee gpurun_out/r02u_bench.log
for c in "16384 32" "65536 8 5"; do timeout 300 python scripts/grid_stats.py $c 2>&1 | grep -E "^C5|sweep" | tee -a gpurun_out/r02u_grid_stats.log; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02u_launches.csv python bench.py --steps 3 --warmup 3 > /dev/null 2>&1
python - <<'PY'
import csv,io,collections
lines=[l for l in open('gpurun_out/r02u_launches.csv') if not l.startswith('==')]
tot=collections.defaultdict(lambda:[0,0.0])
for r in csv.DictReader(io.StringIO(''.join(lines))):
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(r['Metric Value'].replace(',','')); u=r.get('Metric Unit','ns'); v*={'ns':1e-3,'us':1.0,'ms':1e3}.get(u,1e-3)
    t=tot[r['Kernel Name'].split('(')[0]]; t[0]+=1; t[1]+=v
for k,(n,v) in sorted(tot.items(), key=lambda kv:-kv[1][1]): print(f'{k[:60]:60s} {n:4d} {v/n:9.1f} us')
PY
