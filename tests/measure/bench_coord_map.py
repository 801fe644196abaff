"""Measurement of the motion-distance map: aurdf_coord_dist_map on the pose track of
a whole run vs (a) the vectorised numpy oracle and (b) the reference's loop structure
(coord_map.py:253-286: one rotation call per (step, j, k) element) timed on a bounded sample."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from autourdf_b200.coord_map import coord_dist_map
from oracle import coord_map_oracle as C
from test_coord_map import _track

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = []
for T, K in ((50, 20), (50, 30), (50, 48), (100, 128)):
    M = _track(T, K, seed=T + K)
    Md = torch.as_tensor(M).cuda()
    for _ in range(3):
        m, s = coord_dist_map(Md, 0.9)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m, s = coord_dist_map(Md, 0.9)
    e1.record(); torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter(); mo, so = C.coord_dist_map(M, 0.9); np_ms = (time.perf_counter() - t0) * 1e3
    err = float(np.abs(m.cpu().numpy() - mo).max())
    # the reference's structure: per element calls, on ONE step, scaled to T-1 steps
    t0 = time.perf_counter()
    rv = C.rotmat_to_rotvec(np.einsum("kji,kjl->kil", M[0, :, :3, :3], M[1, :, :3, :3]))
    td = M[1, :, :3, 3] - M[0, :, :3, 3]
    dx, dr = np.zeros((K, K)), np.zeros((K, K))
    for j in range(K):
        for k in range(K):
            dx[j, k] = np.linalg.norm(td[j] - td[k], ord=2)
            dr[j, k] = C.rotvec_geodesic_distance(torch.tensor(rv[j]).numpy()[None], torch.tensor(rv[k]).numpy()[None])[0]
    for j in range(K):
        for k in range(K):
            _ = np.linalg.norm(dx[j] - dx[k], ord=2) + np.linalg.norm(dr[j] - dr[k], ord=2)
    loop_ms = (time.perf_counter() - t0) * 1e3 * (T - 1)
    rows.append((T, K, gpu_ms, np_ms, loop_ms, err))
    print(rows[-1], flush=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_coord_map.md"), "w") as f:
    f.write(f"# coord_dist_map(diff=True): pairwise cluster motion-distance map of a whole pose track ({tag}, 1x B200, float64)\n\n")
    f.write("CPU columns: the vectorised numpy oracle, and the reference's own loop structure (coord_map.py:253-286, one call per "
            "(step, j, k) element) timed on one step and scaled to T-1 steps.\n\n")
    f.write("| frames T | clusters K | GPU ms (3 kernels) | numpy oracle ms | reference-structured loops ms (est.) | max abs diff vs oracle |\n|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %d | %d | %.3f | %.1f | %.0f | %.1e |\n" % r)
