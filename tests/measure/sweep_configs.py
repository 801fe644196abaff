"""Throughput + parity table over BASELINE.json's configs (GPU): C1-C4 and points of the C5 sweep.
Writes profiles/<tag>_configs.md.   python tests/measure/sweep_configs.py r01"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci
from oracle import icp_oracle as O
from helpers import ill_posed_tiles

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
O.build(); threads = O.use_all_host_threads()
cases = [("C1 wx200", dict(synth.CONFIGS["wx200"])), ("C2 wx200_5", dict(synth.CONFIGS["wx200_5"])),
         ("C3 franka", dict(synth.CONFIGS["franka"])), ("C4 allegro_hand", dict(synth.CONFIGS["allegro_hand"]))]
for n, k in ((1024, 8), (4096, 32), (8192, 8), (16384, 32), (16384, 128), (32768, 16), (65536, 128), (65536, 8)):
    cases.append((f"C5 {n}x{k}", dict(n_points=n, n_clusters=k, n_seq=1, n_frames=6 if n >= 16384 else 11, dof=5, cid=5)))

rows = []
for name, cfg in cases:
    b = synth.make_batch(**cfg)
    d = ci.batch_to_device(b)
    max_src = int(np.diff(b.src_off).max())
    r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"], max_src_per_tile=max_src)
    torch.cuda.synchronize()
    plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], r.needed_capacity() + 64, max_src)
    run = lambda: plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
    run(); torch.cuda.synchronize()
    reps = 20 if b.src.shape[0] < 200000 else 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        o = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    g = dict(T=o.T.cpu().numpy(), corr=o.corr.cpu().numpy(), iters=o.iters.cpu().numpy(), ntgt=o.ntgt.cpu().numpy())
    t0 = time.perf_counter()
    ref = O.masked_icp_sweep(b.src, b.src_off, b.tgt, b.tgt_off, b.tile_frame, b.box, b.box_off, b.init_T, use_kdtree=True)
    cpu_s = time.perf_counter() - t0
    ill = ill_posed_tiles(b, ref)        # reported only: rank-deficient tiles are compared like every other tile
    keep = np.ones(b.n_tiles, bool)
    pk = np.ones(b.src.shape[0], bool)
    corr_bad = int(((g["corr"] != ref["corr"]) & pk).sum())
    it_bad = int(((g["iters"] != ref["iters"]) & keep).sum())
    perr = float(np.abs(g["T"][keep] - ref["T"][keep]).max())
    ns = np.diff(b.src_off)
    rows.append((name, b.n_frames, b.n_tiles, int(ns.mean()), int(g["ntgt"].mean()), float(ref["iters"].mean()), int(ref["iters"].max()),
                 ms, b.n_frames / (ms * 1e-3), b.n_frames / cpu_s, corr_bad, it_bad, perr, len(ill)))
    print(rows[-1], flush=True)

with open(os.path.join(ROOT, "profiles", f"{tag}_configs.md"), "w") as f:
    f.write(f"# Cluster-ICP sweep over BASELINE.json configs ({tag}, 1x B200, device-resident, float64)\n\n")
    f.write(f"CPU column: oracle/icp_oracle.c, k-d tree, OpenMP over tiles on {threads} host threads (the strongest CPU port, not the reference's structure).\n")
    f.write("Parity columns are against that oracle: mismatching correspondence indices / iteration counts over EVERY tile (rank-deficient ones included: strict pose fit, DESIGN.md) and max pose error.\n\n")
    f.write("| config | frames | tiles | mean n_s | mean n_t | mean / max ICP iters | GPU ms | GPU frames/s | CPU frames/s | corr mismatches | iter mismatches | max pose err | rank-deficient tiles (compared too) |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r_ in rows:
        f.write("| %s | %d | %d | %d | %d | %.1f / %d | %.3f | %.0f | %.0f | %d | %d | %.1e | %d |\n" % r_)
