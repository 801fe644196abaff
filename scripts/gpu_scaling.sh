#!/bin/bash
# Multi-GPU bench under torchrun on one box: weak scaling (the driver's contract) and strong scaling of one
# batch (allegro_hand C4, a C5 sweep point).  usage: scripts/gpu_scaling.sh <N> [tag]
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
run() {   # name, extra args...
  name=$1; shift
  t0=$(date +%s.%N)
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 50 --warmup 5 "$@" \
      > gpurun_out/${TAG}_scale_${name}_$N.json 2> gpurun_out/${TAG}_scale_${name}_$N.err
  rc=$?
  echo "$name N=$N rc=$rc wall=$(python -c "import time; print(round(time.time()-$t0,1))")s"
  grep '^{' gpurun_out/${TAG}_scale_${name}_$N.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d['detail'])"
}
run weak
AURDF_BENCH_RANK_SEEDS=1 run weak_rank_seeds
run strong_allegro --scaling strong --workload allegro_hand
run strong_c5 --scaling strong --workload c5:16384x128x40
NCCL_DEBUG=INFO run weak_nccl_info
