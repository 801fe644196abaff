"""Measurement of the chamfer-L1 operator at the reference's sizes: forward
(both NN directions + reductions) and forward+backward, CUDA events, vs a torch-CPU cdist(p=1)
baseline on all host threads (the SURVEY's proxy for pytorch3d's CPU path).  Appends a table to
profiles/<tag>_chamfer.md."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from autourdf_b200.chamfer import chamfer_distance

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
torch.set_num_threads(os.cpu_count())
rows = []
for P in (1024, 2048, 5000, 10000, 20000):
    rng = np.random.default_rng(P)
    x = torch.from_numpy(rng.normal(0, 0.2, (1, P, 3)).astype(np.float32)).cuda().requires_grad_(True)
    y = torch.from_numpy(rng.normal(0, 0.2, (1, P, 3)).astype(np.float32)).cuda()

    def fwd():
        return chamfer_distance(x, y, norm=1)[0]

    def fwdbwd():
        x.grad = None
        chamfer_distance(x, y, norm=1)[0].backward()

    res = {}
    for name, fn in (("fwd", fwd), ("fwd+bwd", fwdbwd)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps * 1e3      # us
    xc, yc = x.detach().cpu().requires_grad_(True), y.cpu()

    def cpu():
        xc.grad = None
        d = torch.cdist(xc, yc, p=1)
        (d.min(2).values.mean() + d.min(1).values.mean()).backward()

    cpu()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 2.0:
        cpu(); n += 1
    cpu_us = (time.perf_counter() - t0) / n * 1e6
    pairs = 2.0 * P * P
    rows.append((P, res["fwd"], res["fwd+bwd"], pairs / (res["fwd"] * 1e-6) / 1e9, cpu_us, cpu_us / res["fwd+bwd"]))
    print(rows[-1], flush=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_chamfer.md"), "w") as f:
    f.write(f"# chamfer_distance(x, y, norm=1), x and y of P points each ({tag}, 1x B200, float32)\n\n")
    f.write(f"CPU column: torch.cdist(p=1) + min + backward on {os.cpu_count()} host threads (proxy for pytorch3d's CPU knn).\n\n")
    f.write("| P | GPU fwd us | GPU fwd+bwd us | G pair-evals/s (fwd) | CPU fwd+bwd us | CPU/GPU |\n|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %d | %.1f | %.1f | %.1f | %.0f | %.0fx |\n" % r)
