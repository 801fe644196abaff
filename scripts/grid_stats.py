"""Search statistics of icp_grid_kernel on one C5 sweep point (GPU): queries, cache hits, block scans, candidates per
scan, exact fallbacks.   python scripts/grid_stats.py 16384 32 [frames]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci

n, k = int(sys.argv[1]), int(sys.argv[2])
fr = int(sys.argv[3]) if len(sys.argv) > 3 else 6
b = synth.make_batch(n_points=n, n_clusters=k, n_seq=1, n_frames=fr, dof=5, cid=5)
d = ci.batch_to_device(b)
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autourdf_b200", "libaurdf.so"))
lib.aurdf_debug_set_clock_buffer.argtypes = [C.c_void_p]
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
run = lambda: ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"],
                           max_src_per_tile=int(np.diff(b.src_off).max()))
r = run(); torch.cuda.synchronize()
lib.aurdf_debug_set_clock_buffer(buf.data_ptr())
buf[15] = 1                                   # counters
r = run(); torch.cuda.synchronize()
c = buf.cpu().numpy()[:16].copy()
buf.zero_(); buf[15] = 2                      # phase timers (separately: the counters' atomics distort them)
r = run(); torch.cuda.synchronize()
c[6:15] = buf.cpu().numpy()[6:15]
lib.aurdf_debug_set_clock_buffer(None)
it = r.iters.cpu().numpy(); nt = r.ntgt.cpu().numpy(); ns = np.diff(b.src_off)
print(f"C5 {n}x{k}: {b.n_tiles} tiles, mean n_s {ns.mean():.0f}, mean n_t {nt.mean():.0f}, iters mean {it.mean():.1f} max {it.max()}")
print(f"queries {c[0]}, cache hits {c[1]} ({100*c[1]/max(c[0],1):.1f} %), block scans {c[2]} ({c[2]/max(c[0]-c[1],1):.2f} per miss), "
      f"candidates per scan {c[3]/max(c[2],1):.1f} (brute force: {nt.mean():.0f}), exact fallbacks {c[4]} ({100*c[4]/max(c[0],1):.3f} %), "
      f"whole-grid scans {c[5]} ({100*c[5]/max(c[2],1):.2f} %)")
passes = max(c[7], 1)
print(f"per pass (thread 0 of every CTA, cycles): wait at barrier A {c[12]/passes:.0f}, fit + barrier B {c[13]/passes:.0f}")
print(f"longest ICP loop of a CTA: {c[14]} cycles = {c[14]/max(int(it.max())+1,1):.0f} per pass if it is the {it.max()}-iteration tile")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], r.needed_capacity() + 64, int(ns.max()))
f = lambda: plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
f(); torch.cuda.synchronize(); e0.record()
for _ in range(5): f()
e1.record(); torch.cuda.synchronize()
print(f"sweep {e0.elapsed_time(e1)/5:.3f} ms")
