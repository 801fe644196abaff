#!/bin/bash
# compute-sanitizer pass over a small sweep (run on a B200 under gpurun; memcheck ~1 min, racecheck slower):
#   bash scripts/sanitize.sh memcheck|racecheck|synccheck
# The target runs the wx200 config (90 tiles: both ICP kernels' launch paths, host path included) once.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
cat > /tmp/aurdf_sanitize_target.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci
b = synth.make_config("wx200", n_frames=3)
d = ci.batch_to_device(b)
r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"],
                 max_src_per_tile=int(np.diff(b.src_off).max()))
torch.cuda.synchronize()
K = b.n_clusters
ci.masked_icp([b.src[b.src_off[t]:b.src_off[t + 1]] for t in range(K)], [b.box[b.box_off[t]:b.box_off[t + 1]] for t in range(K)],
              b.tgt[b.tgt_off[0]:b.tgt_off[1]], b.init_T[:K].astype(np.float32))
print("iters", r.iters.cpu().numpy()[:10])
PY
timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool "$TOOL" --log-file gpurun_out/sanitize_$TOOL.log \
    python /tmp/aurdf_sanitize_target.py
tail -5 gpurun_out/sanitize_$TOOL.log
