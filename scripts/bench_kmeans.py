"""Measurement of cluster re-sampling: aurdf_resample_clusters over all frames of a
config in ONE launch vs sklearn.cluster.k_means + numpy inverse transforms per frame on the host."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sklearn.cluster import k_means
from autourdf_b200 import synth, _lib

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
L = _lib.lib()
rows = []
for name in ("wx200_5", "franka", "allegro_hand"):
    b = synth.make_config(name)
    K, F, N = b.n_clusters, b.n_frames, b.tgt.shape[0]
    dev = torch.device("cuda")
    cloud = torch.from_numpy(b.tgt).to(dev); off = torch.from_numpy(b.tgt_off).to(dev)
    mats = torch.from_numpy(np.ascontiguousarray(b.init_T)).to(dev)
    labels = torch.empty(N, dtype=torch.int32, device=dev); centers = torch.empty((F, K, 3), dtype=torch.float64, device=dev)
    local = torch.empty((N, 3), dtype=torch.float64, device=dev); loff = torch.empty(F * (K + 1), dtype=torch.int32, device=dev)
    nit = torch.empty(F, dtype=torch.int32, device=dev); inertia = torch.empty(F, dtype=torch.float64, device=dev)
    maxn = int(np.diff(b.tgt_off).max())
    run = lambda: _lib.check(L.aurdf_resample_clusters(_lib.ptr(cloud), _lib.ptr(off), _lib.ptr(mats), F, K, maxn, 300, 1e-4,
                                                       _lib.ptr(labels), _lib.ptr(centers), _lib.ptr(local), _lib.ptr(loff),
                                                       _lib.ptr(nit), _lib.ptr(inertia), _lib.current_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    same = 0
    lab = labels.cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in range(F):
            c = b.tgt[b.tgt_off[f]:b.tgt_off[f + 1]]
            m = b.init_T[f * K:(f + 1) * K]
            _, l, _ = k_means(c, n_clusters=K, init=m[:, :3, 3], n_init=1)
            for k in range(K):
                inv = np.linalg.inv(m[k]); pts = c[l == k]
                _ = (inv @ np.hstack([pts, np.ones((pts.shape[0], 1))]).T)[:3].T
            same += int(np.array_equal(l, lab[b.tgt_off[f]:b.tgt_off[f + 1]]))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    rows.append((name, F, N // F, K, float(nit.float().mean()), ms, F / (ms * 1e-3), cpu_ms, F / (cpu_ms * 1e-3), same))
    print(rows[-1], flush=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_kmeans.md"), "w") as f:
    f.write(f"# resample_cluster (seeded Lloyd k-means + local frames), all frames of a config in one launch ({tag}, 1x B200, float64)\n\n")
    f.write("CPU column: sklearn.cluster.k_means(init=origins, n_init=1) + numpy inverse transforms, frame by frame, as mlp_reg.py:172-217 does.\n\n")
    f.write("| config | frames | points/frame | K | mean Lloyd iters | GPU ms | GPU frames/s | CPU ms | CPU frames/s | frames with labels identical to sklearn |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write("| %s | %d | %d | %d | %.1f | %.3f | %.0f | %.0f | %.0f | %d / %d |\n" % (r + (r[1],)))
