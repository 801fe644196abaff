"""ncu target (GPU): a few chamfer_distance forward + backward calls at the reference's size (P = 5000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200.chamfer import chamfer_distance

rng = np.random.default_rng(0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
x = torch.from_numpy(rng.normal(0, 0.2, (1, P, 3)).astype(np.float32)).cuda().requires_grad_(True)
y = torch.from_numpy(rng.normal(0, 0.2, (1, P, 3)).astype(np.float32)).cuda()
for _ in range(6):
    x.grad = None
    chamfer_distance(x, y, norm=1)[0].backward()
torch.cuda.synchronize()
print("ok")
