"""Sweep time (GPU) of the per-rank wx200_5 workloads of bench.py --gpus N, on one GPU: shows how much the
max over ranks differs from rank 0's time.  python scripts/ab_ranks.py 0 2 4 5 7"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from autourdf_b200 import synth, cluster_icp as ci

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rank in [int(a) for a in sys.argv[1:]] or range(8):
    cfg = dict(synth.CONFIGS["wx200_5"])
    b = synth.make_batch(**cfg, seed=cfg["cid"] * 1000 + 17 * rank)
    d = ci.batch_to_device(b)
    max_src = int(np.diff(b.src_off).max())
    r0 = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"], max_src_per_tile=max_src)
    torch.cuda.synchronize()
    plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], r0.needed_capacity() + 64, max_src)
    run = lambda: plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
    for _ in range(3):
        run()
    ts = []
    for i in range(15):
        flush.fill_(i & 0xFF)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = run(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    it = r.iters.cpu().numpy()
    print(f"SMALL={os.environ.get('AURDF_ICP_SMALL', '1')} rank {rank}: sweep median {np.median(ts) * 1e3:7.1f} us  max n_s {max_src}  iters max {it.max()}", flush=True)
