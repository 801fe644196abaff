#!/bin/bash
# bench.py at N = 1, 2, 4, 8 GPUs of one box (weak scaling as the driver runs it), plus at 8 GPUs: per-rank seeds,
# strong scaling of allegro_hand and of a C5 point, NCCL_DEBUG=INFO.   usage: scripts/gpu_scaling_all.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
one() {   # N name extra-args...
  N=$1; name=$2; shift 2
  t0=$(date +%s.%N)
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 "$@" > gpurun_out/${TAG}_scale_${name}_$N.json 2> gpurun_out/${TAG}_scale_${name}_$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 50 --warmup 5 "$@" \
      > gpurun_out/${TAG}_scale_${name}_$N.json 2> gpurun_out/${TAG}_scale_${name}_$N.err
  fi
  rc=$?
  echo "$name N=$N rc=$rc wall=$(python -c "import time; print(round(time.time()-$t0,1))")s"
  grep '^{' gpurun_out/${TAG}_scale_${name}_$N.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, round(d['e2e']['value']), d['detail'])"
}
export AURDF_BENCH_SKIP_CPU=1
for N in 1 2 4 8; do one $N weak; done
for N in 1 8; do AURDF_BENCH_RANK_SEEDS=1 one $N weak_rank_seeds; done
for N in 1 2 4 8; do one $N strong_allegro --scaling strong --workload allegro_hand; done
for N in 1 8; do one $N strong_c5 --scaling strong --workload c5:16384x128x40; done
NCCL_DEBUG=INFO one 8 weak_nccl_info
grep -c "NCCL INFO" gpurun_out/${TAG}_scale_weak_nccl_info_8.err; grep -m3 -E "nranks|NVLS|Connected all" gpurun_out/${TAG}_scale_weak_nccl_info_8.err
