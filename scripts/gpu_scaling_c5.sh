#!/bin/bash
# strong scaling of a LARGE sweep (C5 65536x8, 40 transitions = 320 tiles of 8192 points): 1 vs 8 GPUs.
TAG=${1:-r02}
mkdir -p gpurun_out
export AURDF_BENCH_SKIP_CPU=1
one() { N=$1; wl=$2
  t0=$(date +%s.%N)
  if [ "$N" = 1 ]; then timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling strong --workload $wl > gpurun_out/${TAG}_scale_big_${wl//:/_}_$N.json 2> gpurun_out/${TAG}_scale_big_$N.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --workload $wl > gpurun_out/${TAG}_scale_big_${wl//:/_}_$N.json 2> gpurun_out/${TAG}_scale_big_$N.err; fi
  echo "$wl N=$N rc=$? wall=$(python -c "import time; print(round(time.time()-$t0,1))")s"
  grep '^{' gpurun_out/${TAG}_scale_big_${wl//:/_}_$N.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, round(d['e2e']['value']), d['detail'])"
}
one 1 c5:65536x8x40
one 8 c5:65536x8x40
one 1 c5:32768x16x40
one 8 c5:32768x16x40
