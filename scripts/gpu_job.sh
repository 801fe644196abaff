mkdir -p gpurun_out
export AURDF_BENCH_SKIP_CPU=1
for shard in frames frames_rr; do for wl in c5:32768x16x40 c5:65536x8x40; do
AURDF_BENCH_SHARD=$shard timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --workload $wl 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$shard $wl', {k: d[k] for k in ('value','ms_per_step','n_gpus')}, d['detail']['sharded_equals_single_gpu'], d['detail']['rank0_tiles'])"
done; done | tee gpurun_out/r02_shard_rr_2.log
