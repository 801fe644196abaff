"""ncu target (GPU): a few launches of the sweep on (a) the slowest wx200_5 tile alone, (b) all 900 tiles.
   python scripts/profile_tiles.py single|all"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci
from autourdf_b200.synth import SweepBatch

b = synth.make_config("wx200_5")
if sys.argv[1:] == ["single"]:
    t = 761   # 68 ICP iterations (scripts/tile_latency.py)
    f = int(b.tile_frame[t])
    b = SweepBatch(b.src[b.src_off[t]:b.src_off[t+1]].copy(), np.array([0, b.src_off[t+1]-b.src_off[t]], np.int32),
                   b.tgt[b.tgt_off[f]:b.tgt_off[f+1]].copy(), np.array([0, b.tgt_off[f+1]-b.tgt_off[f]], np.int32),
                   np.zeros(1, np.int32), b.box[b.box_off[t]:b.box_off[t+1]].copy(),
                   np.array([0, b.box_off[t+1]-b.box_off[t]], np.int32), b.init_T[t:t+1].copy(), 1, 1)
d = ci.batch_to_device(b)
plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], int(b.n_tiles * 400 + 4096), int(np.diff(b.src_off).max()))
for _ in range(4):
    r = plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
torch.cuda.synchronize()
print("iters max", int(r.iters.max()), "tiles", b.n_tiles)
