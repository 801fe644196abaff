"""Shared helpers for the parity tests: run a SweepBatch through the oracle / the CUDA path."""
import numpy as np


def oracle_sweep(O, b, **kw):
    return O.masked_icp_sweep(b.src, b.src_off, b.tgt, b.tgt_off, b.tile_frame, b.box, b.box_off, b.init_T, **kw)


def cuda_sweep(b, pts_dtype=None, **kw):
    import torch
    from autourdf_b200 import cluster_icp as ci
    d = ci.batch_to_device(b, pts_dtype=pts_dtype or torch.float64)
    r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"],
                     d["init_T"], max_src_per_tile=int(np.diff(b.src_off).max()) if b.n_tiles else 0, **kw)
    torch.cuda.synchronize()
    return dict(T=r.T.cpu().numpy(), world=r.world.cpu().numpy(), corr=r.corr.cpu().numpy(),
                fitness=r.fitness.cpu().numpy(), rmse=r.rmse.cpu().numpy(), iters=r.iters.cpu().numpy(),
                ntgt=r.ntgt.cpu().numpy())


def assert_parity(g, o, pose_tol=1e-5, what=""):
    """bit-exact correspondence indices, masked counts and iteration counts; poses within 1e-5
    (north_star tolerance); world points within 1e-5"""
    assert np.array_equal(g["ntgt"], o["ntgt"]), f"{what}: masked target counts differ"
    bad = np.nonzero(g["corr"] != o["corr"])[0]
    assert bad.size == 0, f"{what}: {bad.size} correspondence indices differ, first at {bad[:5]}"
    assert np.array_equal(g["iters"], o["iters"]), f"{what}: iteration counts differ at {np.nonzero(g['iters'] != o['iters'])[0][:5]}"
    assert np.abs(g["T"] - o["T"]).max() <= pose_tol, f"{what}: pose error {np.abs(g['T'] - o['T']).max()}"
    if g["world"].size:
        assert np.abs(g["world"] - o["world"]).max() <= pose_tol
    assert np.abs(g["fitness"] - o["fitness"]).max() <= 1e-12
    assert np.abs(g["rmse"] - o["rmse"]).max() <= 1e-9
