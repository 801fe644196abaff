#!/bin/bash
# compute-sanitizer pass over a small sweep (run on a B200 under gpurun; memcheck ~1 min, racecheck slower):
#   bash scripts/sanitize.sh memcheck|racecheck|synccheck
# The target runs the wx200 config (small-tile kernel, host path included), one tile through each variant of the
# grid-search kernel, and the fused chamfer forward / backward.
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
# the grid-search kernel: one-CTA variant (a 500 x 900 tile beside small ones) and 8-CTA cluster variant (1100 x 1500),
# a few iterations each; then the fused chamfer forward / backward
rng = np.random.default_rng(1)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for ns, nt, it in ((500, 900, 4), (1100, 1500, 3)):
    tgt = rng.uniform(-0.1, 0.1, size=(nt, 3)); tgt[:, 2] *= 0.2
    src = tgt[rng.integers(0, nt, size=ns)] + rng.normal(scale=2e-3, size=(ns, 3))
    small = tgt[:60] + 1e-3
    s_off = np.array([0, ns, ns + 60], np.int32); t_off = np.array([0, nt, nt + nt], np.int32)
    r = ci.icp_sweep(t(np.concatenate([src, small])), t(s_off), t(np.concatenate([tgt, tgt])), t(t_off), t(np.arange(2, dtype=np.int32)),
                     None, None, t(np.stack([np.eye(4)] * 2)), max_src_per_tile=ns, max_iter=it)
    torch.cuda.synchronize()
    print("grid", ns, nt, r.iters.cpu().numpy())
from autourdf_b200.chamfer import chamfer_distance
x = torch.from_numpy(rng.normal(0, 0.2, (2, 700, 3)).astype(np.float32)).cuda().requires_grad_(True)
y = torch.from_numpy(rng.normal(0, 0.2, (2, 900, 3)).astype(np.float32)).cuda()
for _ in range(2):
    chamfer_distance(x, y, norm=1)[0].backward()
torch.cuda.synchronize()
print("chamfer ok")
PY
timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool "$TOOL" --log-file gpurun_out/sanitize_$TOOL.log \
    python /tmp/aurdf_sanitize_target.py
tail -5 gpurun_out/sanitize_$TOOL.log
