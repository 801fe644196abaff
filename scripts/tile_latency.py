"""Microbenchmark (GPU): per-iteration latency of one tile running alone vs the full launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci
from autourdf_b200.synth import SweepBatch

b = synth.make_config("wx200_5")
d = ci.batch_to_device(b)
r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"],
                 max_src_per_tile=int(np.diff(b.src_off).max()))
iters = r.iters.cpu().numpy(); ntgt = r.ntgt.cpu().numpy()
t = int(os.environ.get("AURDF_TILE", iters.argmax()))
print("tile", t, "iters", iters[t], "ns", b.src_off[t+1]-b.src_off[t], "nt", ntgt[t])

def single(t):
    f = int(b.tile_frame[t])
    return SweepBatch(b.src[b.src_off[t]:b.src_off[t+1]].copy(), np.array([0, b.src_off[t+1]-b.src_off[t]], np.int32),
                      b.tgt[b.tgt_off[f]:b.tgt_off[f+1]].copy(), np.array([0, b.tgt_off[f+1]-b.tgt_off[f]], np.int32),
                      np.zeros(1, np.int32), b.box[b.box_off[t]:b.box_off[t+1]].copy(),
                      np.array([0, b.box_off[t+1]-b.box_off[t]], np.int32), b.init_T[t:t+1].copy(), 1, 1)

def timeit(batch, max_iter=10000, reps=20):
    dd = ci.batch_to_device(batch)
    plan = ci.IcpSweep(batch.n_tiles, batch.src.shape[0], int(batch.n_tiles * (np.diff(batch.tgt_off).max() + 2)),
                       int(np.diff(batch.src_off).max()))
    L = plan.lib
    for _ in range(3):
        plan.run(dd["src"], dd["src_off"], dd["tgt"], dd["tgt_off"], dd["tile_frame"], dd["box"], dd["box_off"], dd["init_T"], max_iter=max_iter)
    torch.cuda.synchronize()
    import ctypes as C
    L.aurdf_icp_profile_enable(1)
    for _ in range(reps):
        o = plan.run(dd["src"], dd["src_off"], dd["tgt"], dd["tgt_off"], dd["tile_frame"], dd["box"], dd["box_off"], dd["init_T"], max_iter=max_iter)
    torch.cuda.synchronize()
    L.aurdf_icp_profile_enable(0)
    ms, n = C.c_double(), C.c_int32()
    L.aurdf_icp_profile_collect(C.byref(ms), C.byref(n))
    return ms.value / n.value * 1e3, o.iters.cpu().numpy()

one = single(t)
for mi in (0, 1, 2, 4, 8, 16, 32, 10000):
    us, it = timeit(one, mi)
    print(f"single tile max_iter={mi:5d}: {us:8.1f} us  iters={it[0]}")
for mi in (0, 1, 2, 4, 8, 16, 10000):
    us, it = timeit(b, mi)
    print(f"all 900 tiles max_iter={mi:5d}: {us:8.1f} us  mean iters={it.mean():.1f}")

# ---- per-phase cycle stamps of the single tile (debug hook) ----
import ctypes as C
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autourdf_b200", "libaurdf.so"))
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.aurdf_debug_set_clock_buffer.argtypes = [C.c_void_p]
lib.aurdf_debug_set_clock_buffer(buf.data_ptr())
timeit(one, 10000, reps=1)
lib.aurdf_debug_set_clock_buffer(None)
c = buf.cpu().numpy()
if os.environ.get("AURDF_ICP_SMALL", "128") != "0":
    # icp_small_kernel: (clock, id) pairs of thread 0 of tile 0
    ids = {0: "iteration start", 1: "P update", 2: "float32 scan", 3: "merge+certificate+exact distance", 4: "exact rescan branch",
           5: "pass end", 6: "barrier 1", 7: "moment reduction", 8: "barrier 2", 20: "totals+covariance", 21: "rotation fit",
           9: "translation+store U (fit done)"}
    if os.environ.get("AURDF_ICP_SMALL", "2") == "2":   # icp_small2_kernel
        ids = {0: "iteration start", 1: "P update", 2: "cache test", 3: "float32 scan (misses)", 4: "merge+certificate+exact distance",
               5: "pass end", 7: "moment reduction", 8: "barrier A", 20: "totals+covariance", 21: "rotation fit",
               9: "translation+store U (fit done)", 22: "strict: ordered sums", 23: "strict: Jacobi SVD + pose",
               30: "reduce: loads issued", 31: "reduce: products done"}
    c = c.reshape(-1, 2)
    c = c[c[:, 0] > 0]
    from collections import OrderedDict
    d = OrderedDict()
    for k in range(1, len(c)):
        d.setdefault((int(c[k - 1, 1]), int(c[k, 1])), []).append(int(c[k, 0] - c[k - 1, 0]))
    for (a_, b_), v in d.items():
        print("%-22s -> %-34s: median %6d  p90 %6d  n %d" % (ids.get(a_, a_), ids.get(b_, b_), np.median(v), np.percentile(v, 90), len(v)))
    it0 = c[c[:, 1] == 0, 0]
    print("whole iteration: median %d cycles" % np.median(np.diff(it0)))
    sys.exit(0)
c = c[c > 0]
first, rest = c[0], c[1:]
k = (len(rest)) // 7
st = rest[:7 * k].reshape(k, 7)
print("iterations stamped", k)
if os.environ.get("AURDF_ICP_SMALL", "128") == "0":
    names = ["compose + P update", "NN scan", "S-merge + moment sums", "warp_sum16 + store (pass done)", "barrier A wait", "totals + pose fit (lane 0)"]
else:   # icp_small_kernel
    names = ["compose + P update", "NN scan + merge + exact distance", "barrier 1 wait", "moment reduction", "barrier 2 wait", "pose fit (lane 0)"]
for i, nme in enumerate(names):
    print("%-34s: median %d cycles" % (nme, np.median(st[:, i + 1] - st[:, i])))
print("%-34s: median %d" % ("barrier B + loop overhead", np.median(st[1:, 0] - st[:-1, 6])))
print("%-34s: median %d" % ("whole iteration", np.median(st[1:, 0] - st[:-1, 0])))
