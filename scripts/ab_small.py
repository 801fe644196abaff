"""A/B (GPU): sweep time of the named configs under the current AURDF_ICP_SMALL setting."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in sys.argv[1:] or ["wx200_5", "franka", "allegro_hand"]:
    b = synth.make_config(name)
    d = ci.batch_to_device(b)
    max_src = int(np.diff(b.src_off).max())
    r0 = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"],
                      max_src_per_tile=max_src)
    torch.cuda.synchronize()
    plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], r0.needed_capacity() + 64, max_src)
    L = plan.lib
    run = lambda: plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    L.aurdf_icp_profile_enable(1)
    ts = []
    for i in range(30):
        flush.fill_(i & 0xFF)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = run(); e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    L.aurdf_icp_profile_enable(0)
    ms, n = C.c_double(), C.c_int32()
    L.aurdf_icp_profile_collect(C.byref(ms), C.byref(n))
    it = r.iters.cpu().numpy()
    same = bool(np.array_equal(it, r0.iters.cpu().numpy()))
    print(f"SMALL={os.environ.get('AURDF_ICP_SMALL', '1')} MINB={os.environ.get('AURDF_ICP_SMALL_MINB', '6')} {name:13s} tiles {b.n_tiles:5d}  sweep median {np.median(ts)*1e3:8.1f} us"
          f"  min {min(ts)*1e3:8.1f} us  icp kernels {ms.value / n.value * 1e3:8.1f} us  iters mean {it.mean():.2f} max {it.max()}  "
          f"frames/s {b.n_frames / (np.median(ts) * 1e-3):9.0f}", flush=True)
