"""ncu target (GPU): a few sweeps of one C5 point.   python scripts/profile_c5.py 16384 32 [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci

n, k = int(sys.argv[1]), int(sys.argv[2])
fr = int(sys.argv[3]) if len(sys.argv) > 3 else 6
b = synth.make_batch(n_points=n, n_clusters=k, n_seq=1, n_frames=fr, dof=5, cid=5)
d = ci.batch_to_device(b)
ms = int(np.diff(b.src_off).max())
r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"], max_src_per_tile=ms)
torch.cuda.synchronize()
plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], r.needed_capacity() + 64, ms)
for _ in range(3):
    r = plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
torch.cuda.synchronize()
print("iters max", int(r.iters.max()), "tiles", b.n_tiles)
