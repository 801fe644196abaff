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
    # the same two calls replayed from a CUDA graph: what the operator costs on the device once the Python / launch
    # overhead of the eager call (ctypes, autograd.Function, allocator) is out of the way
    for name, fn in (("fwd graph", fwd), ("fwd+bwd graph", fwdbwd)):
        x.grad = torch.zeros_like(x)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                (fwd() if name == "fwd graph" else chamfer_distance(x, y, norm=1)[0].backward())
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if name == "fwd graph":
                out = fwd()
            else:
                chamfer_distance(x, y, norm=1)[0].backward()
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps * 1e3
        del g
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
    rows.append((P, res["fwd"], res["fwd+bwd"], res["fwd graph"], res["fwd+bwd graph"], pairs / (res["fwd graph"] * 1e-6) / 1e9,
                 cpu_us, cpu_us / res["fwd+bwd"]))
    print(rows[-1], flush=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_chamfer.md"), "w") as f:
    f.write(f"# chamfer_distance(x, y, norm=1), x and y of P points each ({tag}, 1x B200, float32)\n\n")
    f.write(f"CPU column: torch.cdist(p=1) + min + backward on {os.cpu_count()} host threads (proxy for pytorch3d's CPU knn).\n\n")
    f.write("Eager = the Python call as the reference's loop issues it (one fused launch forward, memset + one launch backward); graph = the same "
            "calls replayed from a CUDA graph (device cost without the per-call Python / launch overhead).  Issue roofline of the forward kernel: "
            "~7 lane instructions per (query, target) pair (5 FADD + min tree + loop), peak 148 SMs x 128 lanes x 1.965 GHz = 37.2 T lane-instr/s "
            "=> 5.3 T pairs/s; the last column is the fraction the graph-replayed forward reaches.\n\n")
    f.write("| P | eager fwd us | eager fwd+bwd us | graph fwd us | graph fwd+bwd us | G pair-evals/s (graph fwd) | fraction of issue roof | CPU fwd+bwd us | CPU / eager GPU |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f | %.0f | %.0fx |\n" % (r[0], r[1], r[2], r[3], r[4], r[5], r[5] * 7 / 37200.0, r[6], r[7]))
