"""Multi-GPU layer of the cluster-ICP sweep: one process per GPU (torch.distributed), tiles
sharded by contiguous frame blocks (batch mode: every target cloud lives on exactly one GPU) or by
contiguous tile ranges (real-pipeline mode: the K clusters of the current frame transition are
split over the GPUs, each holding a replica of the frame's cloud), ONE all-gather of the fitted
poses per sweep.

Why it shards: the K cluster tiles of a frame are independent (the loop at reference
PointCloud/cluster_icp.py:131 carries no state) and so are frames once their init poses are
given; sequences are independent too (mlp_reg.py:434-435).  Nothing is exchanged on the data
path; the only collective is the all-gather of (tiles x 16) float64 poses (+ fitness, rmse,
iteration counts) so every rank ends the sweep holding all poses, as the single-process
reference does.  Correspondences and world clusters stay sharded.

The local compute is injected (``run_local``), so the partition / gather logic is exercised on
CPU with gloo in tests/ and with the CUDA sweep + NCCL on the GPUs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def _to_np(a):
    """numpy view/copy of a numpy array or a (possibly CUDA) torch tensor"""
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


def _require_frame_major(batch):
    """frame blocks are contiguous tile ranges only if tiles are sorted by frame (every producer in this
    package emits them that way); anything else must shard by="tiles" """
    if batch.n_tiles > 1 and not np.all(np.diff(batch.tile_frame) >= 0):
        raise ValueError("sharding by frames needs frame-major tiles (tile_frame non-decreasing); "
                         "use by='tiles' for an unsorted batch")


def frame_partition(batch, world: int):
    """Contiguous frame ranges [f0, f1) per rank, balanced by the pair-evaluation estimate
    sum(n_s * M) of each frame (keeps every target cloud on exactly one GPU)."""
    _require_frame_major(batch)
    F = batch.n_frames
    ns = np.diff(batch.src_off).astype(np.float64)
    M = np.diff(batch.tgt_off).astype(np.float64)
    cost = np.zeros(F)
    np.add.at(cost, batch.tile_frame, ns * M[batch.tile_frame])
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        f = int(np.searchsorted(cum, target, side="left"))
        f = min(max(f, cuts[-1]), F)
        cuts.append(f)
    cuts.append(F)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def frame_partition_interleaved(batch, world: int):
    """Frames dealt round-robin: rank r owns the transitions r, r + world, r + 2 world, ...  ICP work per tile is its
    iteration count, which drifts along a long chain of frames (late frames of the 40-transition C5 chains hold the
    170-190-iteration tiles): contiguous blocks balanced by points then leave one rank with all the slow frames
    (profiles/r02_scaling.md), a round-robin deal gives every rank a sample of the whole chain."""
    _require_frame_major(batch)
    return [np.arange(r, batch.n_frames, world, dtype=np.int64) for r in range(world)]


def tile_partition(batch, world: int):
    """Contiguous tile ranges [b0, b1) per rank, balanced by the pair-evaluation estimate n_s * M of
    each tile.  Frames may be split between ranks (their clouds are then replicated): this is how one
    sweep of K clusters (real-pipeline mode) spreads over the GPUs."""
    B = batch.n_tiles
    ns = np.diff(batch.src_off).astype(np.float64)
    M = np.diff(batch.tgt_off).astype(np.float64)
    cost = ns * M[batch.tile_frame] if B else np.zeros(0)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cum, cum[-1] * r / world, side="left"))
        cuts.append(min(max(b, cuts[-1]), B))
    cuts.append(B)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def sharded_sweep(batch, run_local, group=None, device="cpu", by="frames"):
    """Run ``run_local(sub_batch) -> dict(T (b,4,4), fitness, rmse, iters)`` on this rank's
    share and all-gather the per-tile results.  ``by="frames"``: contiguous frame blocks; ``by="frames_rr"``: frames
    dealt round-robin (balances chains whose difficulty drifts);
    ``by="tiles"``: contiguous tile ranges with replicated target clouds (a single frame's K
    clusters over several GPUs).  Returns (gathered dict over ALL tiles in the original tile
    order, this rank's local result dict, this rank's (f0, f1) or (b0, b1))."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    assert by in ("frames", "tiles", "frames_rr")
    order = None   # tile indices in gathered (rank-major) order, when that is not the original order
    if by == "frames_rr":
        parts = frame_partition_interleaved(batch, world)
        sel = [np.nonzero(np.isin(batch.tile_frame, fr))[0] for fr in parts]
        sub, _ = batch.frame_select(parts[rank])
        counts = [int(t.size) for t in sel]
        order = np.concatenate(sel) if sel else np.zeros(0, dtype=np.int64)
        f0, f1 = (int(parts[rank][0]), int(parts[rank][-1]) + 1) if parts[rank].size else (0, 0)
    elif by == "tiles":
        parts = tile_partition(batch, world)
        f0, f1 = parts[rank]
        sub = batch.tile_slice(f0, f1)
        counts = [b - a for a, b in parts]
    else:
        parts = frame_partition(batch, world)
        f0, f1 = parts[rank]
        sub = batch.frame_slice(f0, f1)
        # tiles per rank (frame-major tile order => each rank owns a contiguous tile range)
        counts = [int(((batch.tile_frame >= a) & (batch.tile_frame < b)).sum()) for a, b in parts]
    local = run_local(sub) if sub.n_tiles else dict(T=np.zeros((0, 4, 4)), fitness=np.zeros(0), rmse=np.zeros(0),
                                                    iters=np.zeros(0, dtype=np.int32))
    if world == 1:
        return dict(T=_to_np(local["T"]), fitness=_to_np(local["fitness"]), rmse=_to_np(local["rmse"]),
                    iters=_to_np(local["iters"])), local, (f0, f1)
    width = max(max(counts), 1)
    # one padded (width, 19) float64 payload per rank: 16 pose entries + fitness + rmse + iters
    pay = torch.zeros((width, 19), dtype=torch.float64, device=device)
    n = counts[rank]
    if n:
        as_t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=device)
        pay[:n, :16] = as_t(local["T"]).reshape(n, 16)
        pay[:n, 16] = as_t(local["fitness"])
        pay[:n, 17] = as_t(local["rmse"])
        pay[:n, 18] = as_t(local["iters"])
    out = torch.empty((world * width, 19), dtype=torch.float64, device=device)   # concatenated layout
    dist.all_gather_into_tensor(out, pay, group=group)
    out = out.cpu().numpy().reshape(world, width, 19)
    rows = np.concatenate([out[r, :counts[r]] for r in range(world)], axis=0)
    if order is not None:   # back to the original tile order
        inv = np.empty_like(rows)
        inv[order] = rows
        rows = inv
    return dict(T=rows[:, :16].reshape(-1, 4, 4).copy(), fitness=rows[:, 16].copy(), rmse=rows[:, 17].copy(),
                iters=rows[:, 18].astype(np.int32)), local, (f0, f1)


def cuda_run_local(device=None, **kw):
    """``run_local`` backed by the CUDA sweep on this rank's GPU."""
    from . import cluster_icp as ci

    def run(sub):
        d = ci.batch_to_device(sub, device=device or torch.device("cuda", torch.cuda.current_device()))
        r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"],
                         d["init_T"], max_src_per_tile=int(np.diff(sub.src_off).max()), **kw)
        return dict(T=r.T, fitness=r.fitness, rmse=r.rmse, iters=r.iters, corr=r.corr, world=r.world, ntgt=r.ntgt)

    return run


class ShardedSweep:
    """Device-resident sharding of ONE batch over the ranks of ``group`` (strong scaling: the total work
    is fixed, every rank owns a contiguous frame block, a contiguous tile range, or -- ``by="frames_rr"`` -- every
    world-th frame).  The partition, the upload of this
    rank's share and the plan (outputs + workspace) happen once; ``run()`` is then this rank's five kernel
    launches plus ONE ``all_gather_into_tensor`` of the fitted poses (tiles x 16 float64, padded to the
    widest share), all on the current stream with no host synchronisation.  Reference: the per-cluster
    loop at PointCloud/cluster_icp.py:131 carries no state between clusters, mlp_reg.py:434-435 none
    between sequences."""

    def __init__(self, batch, device, group=None, by="frames", **sweep_kw):
        from . import cluster_icp as ci
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group, self.kw = group, sweep_kw
        assert by in ("frames", "tiles", "frames_rr")
        self.order = None
        if by == "frames_rr":
            self.parts = frame_partition_interleaved(batch, self.world)
            sel = [np.nonzero(np.isin(batch.tile_frame, fr))[0] for fr in self.parts]
            self.sub, _ = batch.frame_select(self.parts[self.rank])
            self.counts = [int(t.size) for t in sel]
            self.order = np.concatenate(sel)
        elif by == "tiles":
            self.parts = tile_partition(batch, self.world)
            lo, hi = self.parts[self.rank]
            self.sub = batch.tile_slice(lo, hi)
            self.counts = [b - a for a, b in self.parts]
        else:
            self.parts = frame_partition(batch, self.world)
            lo, hi = self.parts[self.rank]
            self.sub = batch.frame_slice(lo, hi)
            self.counts = [int(((batch.tile_frame >= a) & (batch.tile_frame < b)).sum()) for a, b in self.parts]
        self.width = max(max(self.counts), 1)
        n = self.sub.n_tiles
        assert n == self.counts[self.rank]
        dev = torch.device(device)
        self.pay = torch.zeros((self.width, 16), dtype=torch.float64, device=dev)
        self.gathered = torch.empty((self.world * self.width, 16), dtype=torch.float64, device=dev)
        self.plan, self.d = None, None
        if n:
            self.d = ci.batch_to_device(self.sub, device=dev)
            max_src = int(np.diff(self.sub.src_off).max())
            r0 = ci.icp_sweep(self.d["src"], self.d["src_off"], self.d["tgt"], self.d["tgt_off"], self.d["tile_frame"],
                              self.d["box"], self.d["box_off"], self.d["init_T"], max_src_per_tile=max_src, **sweep_kw)
            self.plan = ci.IcpSweep(n, self.sub.src.shape[0], r0.needed_capacity() + 64, max_src, device=dev)
            self.plan.out.T = self.pay[:n].view(n, 4, 4)      # poses land in the all-gather payload directly

    def run(self):
        """one sharded sweep; returns the gathered (world * width, 16) pose buffer (device)"""
        if self.plan is not None:
            d = self.d
            self.plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"],
                          d["init_T"], **self.kw)
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.pay, group=self.group)
        else:
            self.gathered.copy_(self.pay)
        return self.gathered

    def poses(self):
        """(B,4,4) float64 numpy: the gathered poses in the original tile order"""
        g = self.gathered.cpu().numpy().reshape(self.world, self.width, 16)
        rows = np.concatenate([g[r, :self.counts[r]] for r in range(self.world)], axis=0)
        if self.order is not None:   # interleaved frames: back to the original tile order
            inv = np.empty_like(rows)
            inv[self.order] = rows
            rows = inv
        return rows.reshape(-1, 4, 4)
