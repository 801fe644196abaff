import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from autourdf_b200 import synth
from oracle import icp_oracle as O
from helpers import cuda_sweep, oracle_sweep, ill_posed_tiles
b = synth.make_config("franka")
g = cuda_sweep(b); o = oracle_sweep(O, b, use_kdtree=True)
for t in ill_posed_tiles(b, o):
    R, Ro = g["T"][t][:3, :3], o["T"][t][:3, :3]
    print("tile", t, "ntgt", o["ntgt"][t], "iters g/o", g["iters"][t], o["iters"][t], "orth err g/o", np.abs(R @ R.T - np.eye(3)).max(), np.abs(Ro @ Ro.T - np.eye(3)).max(), "det", np.linalg.det(R))
    print(g["T"][t])
