#!/bin/bash
# A/B of the small-tile kernel's residency (CTAs per SM it is compiled for) on the named configs.
# usage: scripts/ab_minb.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
for wl in wx200_5 franka allegro_hand; do
  for mb in 4 5 6; do
    AURDF_BENCH_SKIP_CPU=1 AURDF_ICP_SMALL_MINB=$mb timeout 120 python bench.py --steps 50 --warmup 5 --workload $wl 2>/dev/null | tail -1 > gpurun_out/minb_${wl}_$mb.json
    python -c "
import json; d=json.load(open('gpurun_out/minb_${wl}_$mb.json')); print('$wl minb $mb: value', round(d['value']), ' e2e', round(d['e2e']['value']), ' kernel_ms', round(d['roofline']['kernel_ms'],4))"
  done
done | tee gpurun_out/${TAG}_minb.log
