"""Cluster-ICP sweep: B200 replacement for ``masked_icp`` (reference
PointCloud/cluster_icp.py:118-191, called from PointCloud/mlp_reg.py:325) and for the
open3d ``registration_icp`` call sites (cluster_icp.py:157, link.py:113, Sim/evaluation.py:358).

Two layers:
  * ``IcpSweep`` / ``icp_sweep`` -- packed device tensors in, device tensors out, one launch
    sequence for all (frame, cluster) tiles; this is what the bench and multi-GPU path use;
  * ``masked_icp`` / ``registration_icp`` -- the reference's own signatures (numpy in, numpy
    out), a thin packing layer over the C ABI.
Everything computes in libaurdf.so (hand-written sm_100a kernels); nothing here falls back
to a CPU implementation.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

_TORCH2DT = {torch.float32: _lib.F32, torch.float64: _lib.F64}


@dataclass
class IcpResult:
    T: torch.Tensor          # (B,4,4) f64 fitted poses
    world: torch.Tensor      # (N,3) f64   T applied to the source points
    corr: torch.Tensor       # (N,) i32    index into the tile's frame cloud, -1 = none
    fitness: torch.Tensor    # (B,) f64
    rmse: torch.Tensor       # (B,) f64
    iters: torch.Tensor      # (B,) i32
    ntgt: torch.Tensor       # (B,) i32    masked target points of the tile
    status: torch.Tensor     # (4,) i32    [0]=capacity exceeded, [1..2]=needed capacity lo/hi

    def needed_capacity(self) -> int:
        s = self.status.cpu().numpy().astype(np.int64)
        return int((s[1] & 0xFFFFFFFF) | (s[2] << 32))

    def overflowed(self) -> bool:
        return bool(self.status[0].item())


class IcpSweep:
    """Reusable plan: owns outputs + workspace for a fixed problem shape, launches without
    any host synchronisation (all five kernels go to the current stream)."""

    def __init__(self, n_tiles: int, total_src: int, tgt_capacity: int, max_src_per_tile: int = 0, device=None):
        self.lib = _lib.lib()
        dev = torch.device("cuda") if device is None else torch.device(device)
        self.n_tiles, self.total_src, self.tgt_capacity = int(n_tiles), int(total_src), int(tgt_capacity)
        self.max_src_per_tile = int(max_src_per_tile)
        nbytes = self.lib.aurdf_icp_workspace_bytes(self.n_tiles, self.total_src, self.tgt_capacity)
        self.workspace = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        B, N = self.n_tiles, self.total_src
        self.out = IcpResult(
            T=torch.empty((B, 4, 4), dtype=torch.float64, device=dev),
            world=torch.empty((N, 3), dtype=torch.float64, device=dev),
            corr=torch.empty((N,), dtype=torch.int32, device=dev),
            fitness=torch.empty((B,), dtype=torch.float64, device=dev),
            rmse=torch.empty((B,), dtype=torch.float64, device=dev),
            iters=torch.empty((B,), dtype=torch.int32, device=dev),
            ntgt=torch.empty((B,), dtype=torch.int32, device=dev),
            status=torch.zeros((4,), dtype=torch.int32, device=dev))

    def run(self, src, src_off, tgt, tgt_off, tile_frame, box, box_off, init_T, box_scale=1.2, max_corr=1.0,
            max_iter=10000, rel_fitness=1e-6, rel_rmse=1e-6, ori_only=False) -> IcpResult:
        assert src.dtype == tgt.dtype and src.dtype in _TORCH2DT
        assert init_T.dtype == torch.float64 and init_T.is_contiguous()
        assert (box is None) == (box_off is None)   # box=None: no mask (plain registration_icp)
        box_dt = _lib.F64 if box is None else _TORCH2DT[box.dtype]
        for t in (src, tgt, box, src_off, tgt_off, tile_frame, box_off):
            assert t is None or (t.is_cuda and t.is_contiguous())
        for t in (src_off, tgt_off, tile_frame, box_off):
            assert t is None or t.dtype == torch.int32
        o = self.out
        rc = self.lib.aurdf_icp_sweep(
            _lib.ptr(src), _TORCH2DT[src.dtype], _lib.ptr(src_off), _lib.ptr(tgt), _lib.ptr(tgt_off),
            _lib.ptr(tile_frame), _lib.ptr(box), box_dt, _lib.ptr(box_off), _lib.ptr(init_T),
            self.n_tiles, self.total_src, self.max_src_per_tile, float(box_scale), float(max_corr), int(max_iter),
            float(rel_fitness), float(rel_rmse), int(bool(ori_only)), _lib.ptr(o.T), _lib.ptr(o.world),
            _lib.ptr(o.corr), _lib.ptr(o.fitness), _lib.ptr(o.rmse), _lib.ptr(o.iters), _lib.ptr(o.ntgt),
            _lib.ptr(self.workspace), self.workspace.numel(), self.tgt_capacity, _lib.ptr(o.status),
            _lib.current_stream())
        _lib.check(rc, "aurdf_icp_sweep")
        return o


def icp_sweep(src, src_off, tgt, tgt_off, tile_frame, box, box_off, init_T, tgt_capacity=None,
              max_src_per_tile=0, **kw) -> IcpResult:
    """One-shot sweep on device tensors.  Sizes the compacted-target capacity itself and
    retries once with the exact need if the first guess was too small (this reads the
    4-int status back, i.e. synchronises; use ``IcpSweep`` directly to avoid that)."""
    B, N = int(tile_frame.numel()), int(src.shape[0])
    if tgt_capacity is None:
        tgt_capacity = 4 * N + 2 * B + 1024
    plan = IcpSweep(B, N, tgt_capacity, max_src_per_tile, device=src.device)
    r = plan.run(src, src_off, tgt, tgt_off, tile_frame, box, box_off, init_T, **kw)
    if B and r.overflowed():
        plan = IcpSweep(B, N, r.needed_capacity(), max_src_per_tile, device=src.device)
        r = plan.run(src, src_off, tgt, tgt_off, tile_frame, box, box_off, init_T, **kw)
        assert not r.overflowed()
    return r


def batch_to_device(batch, device="cuda", pts_dtype=torch.float64):
    """SweepBatch (autourdf_b200.synth) -> dict of device tensors in the C-ABI layout."""
    dev = torch.device(device)
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev if dt is None else dev, dtype=dt)
    return dict(src=t(batch.src, pts_dtype), src_off=t(batch.src_off), tgt=t(batch.tgt, pts_dtype),
                tgt_off=t(batch.tgt_off), tile_frame=t(batch.tile_frame), box=t(batch.box),
                box_off=t(batch.box_off), init_T=t(batch.init_T, torch.float64))


def _pack(arrs, dtype):
    """list of (n_k,3) arrays -> packed (sum n_k, 3) array of ``dtype`` + CSR offsets (one concatenate:
    this runs once per call of the drop-in, on the host, inside the caller's frame loop)"""
    off = np.zeros(len(arrs) + 1, dtype=np.int32)
    if not arrs:
        return np.empty((0, 3), dtype=dtype), off
    off[1:] = np.cumsum([a.shape[0] for a in arrs])
    out = np.concatenate(arrs, axis=0, dtype=dtype) if len(arrs) > 1 else np.array(arrs[0], dtype=dtype, order="C")
    return out, off


class HostSweep:
    """Host-buffer path (``aurdf_icp_sweep_host``): numpy in, numpy out.  The context owns streams,
    pinned staging and device buffers.  A single-frame call with pageable arrays (the reference's
    frame loop) is one H2D copy, five kernels, one D2H copy; a large frame-major batch is cut into
    frame blocks on separate streams so copies overlap kernels."""

    def __init__(self, device: int | None = None):
        import ctypes as C
        self.lib = _lib.lib()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        h = C.c_void_p()
        _lib.check(self.lib.aurdf_ctx_create(int(device), C.byref(h)), "aurdf_ctx_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aurdf_ctx_destroy(self._h)
            self._h = None

    __del__ = close

    def copy_bytes(self):
        import ctypes as C
        a, b = C.c_int64(), C.c_int64()
        self.lib.aurdf_ctx_last_copy_bytes(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def run(self, src, src_off, tgt, tgt_off, tile_frame, box, box_off, init_T, box_scale=1.2, max_corr=1.0,
            max_iter=10000, rel_fitness=1e-6, rel_rmse=1e-6, ori_only=False, out=None, want_world=True,
            want_corr=True):
        """All arguments are C-contiguous numpy arrays in the C-ABI layout (box may be None).
        ``want_world`` / ``want_corr`` = False leaves the per-point outputs on the device (``out["world"]`` /
        ``out["corr"]`` are then None): a caller that only needs the fitted poses moves 85 % fewer bytes back."""
        assert src.dtype == tgt.dtype and src.dtype in (np.float32, np.float64)
        assert init_T.dtype == np.float64 and src_off.dtype == np.int32 and tgt_off.dtype == np.int32
        assert tile_frame.dtype == np.int32 and (box is None or box_off.dtype == np.int32)
        for a in (src, tgt, box, src_off, tgt_off, tile_frame, box_off, init_T):
            assert a is None or a.flags.c_contiguous
        B, F, N = int(tile_frame.shape[0]), int(tgt_off.shape[0]) - 1, int(src.shape[0])
        if out is None:
            out = dict(T=np.empty((B, 4, 4)), world=np.empty((N, 3)) if want_world else None,
                       corr=np.empty(N, dtype=np.int32) if want_corr else None,
                       fitness=np.empty(B), rmse=np.empty(B), iters=np.empty(B, dtype=np.int32),
                       ntgt=np.empty(B, dtype=np.int32))
        dt = lambda a: _lib.F32 if a.dtype == np.float32 else _lib.F64
        rc = self.lib.aurdf_icp_sweep_host(
            self._h, _lib.ptr(src), dt(src), _lib.ptr(src_off), _lib.ptr(tgt), _lib.ptr(tgt_off),
            _lib.ptr(tile_frame), _lib.ptr(box), _lib.F64 if box is None else dt(box), _lib.ptr(box_off),
            _lib.ptr(init_T), B, F, float(box_scale), float(max_corr), int(max_iter), float(rel_fitness),
            float(rel_rmse), int(bool(ori_only)), _lib.ptr(out["T"]), _lib.ptr(out["world"] if want_world else None),
            _lib.ptr(out["corr"] if want_corr else None),
            _lib.ptr(out["fitness"]), _lib.ptr(out["rmse"]), _lib.ptr(out["iters"]), _lib.ptr(out["ntgt"]))
        _lib.check(rc, "aurdf_icp_sweep_host")
        return out


_HOST_CTX = {}


def _host_ctx() -> HostSweep:
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if dev not in _HOST_CTX:
        _HOST_CTX[dev] = HostSweep(dev)
    return _HOST_CTX[dev]


def masked_icp(clusters_local, clusters_world, step_pc_np, matrices, visual=False, ori=False, scale=1.2, th=1,
               colors=None, _details=None):
    """Drop-in for the reference ``masked_icp`` (cluster_icp.py:118).

    Args (as upstream): clusters_local list of (n_k,3); clusters_world list of (n_k',3)
    (float32 from mlp_reg.py:121, float64 accepted); step_pc_np (M,3); matrices (K,4,4).
    ``visual`` / ``colors`` drive open3d GUI calls upstream and are accepted and ignored.
    Returns (list of (n_k,3) float64 world clusters, (K,4,4) float64 poses)."""
    K = min(len(clusters_local), len(clusters_world), len(matrices))   # zip() truncation, :131
    as3 = lambda c: c if (type(c) is np.ndarray and c.ndim == 2 and c.shape[1] == 3) else np.asarray(c).reshape(-1, 3)
    cl = [as3(c) for c in clusters_local[:K]]
    cw = [as3(c) for c in clusters_world[:K]]
    for c in cw:
        if c.shape[0] == 0:   # np.min on an empty array, :133
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if not th > 0:
        raise RuntimeError("[Open3D Error] Invalid max_correspondence_distance.")
    if K == 0:
        return [], np.array([])
    wdt = np.float32 if all(c.dtype == np.float32 for c in cw) else np.float64
    src, src_off = _pack(cl, np.float64)
    box, box_off = _pack(cw, wdt)
    tgt = np.ascontiguousarray(step_pc_np, dtype=np.float64).reshape(-1, 3)
    if type(matrices) is np.ndarray and matrices.ndim == 3:
        init = np.array(matrices[:K], dtype=np.float64, order="C")
    else:
        init = np.ascontiguousarray(np.asarray([np.asarray(m, dtype=np.float64) for m in matrices[:K]]).reshape(K, 4, 4))
    r = _host_ctx().run(src, src_off, tgt, np.array([0, tgt.shape[0]], dtype=np.int32), np.zeros(K, dtype=np.int32),
                        box, box_off, init, box_scale=scale, max_corr=th, max_iter=10000, ori_only=ori,
                        want_corr=_details is not None)
    if _details is not None:
        _details.update(r, src_off=src_off)
    # the output buffer is freshly allocated by run(): the per-cluster views own it, no copy needed
    world, o = r["world"], src_off.tolist()
    return [world[o[k]:o[k + 1]] for k in range(K)], r["T"]


@dataclass
class RegistrationResult:
    """Mirror of open3d.pipelines.registration.RegistrationResult"""
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    correspondence_set: np.ndarray
    iterations: int = 0


def registration_icp(source, target, max_correspondence_distance, init=None, max_iteration=30,
                     relative_fitness=1e-6, relative_rmse=1e-6) -> RegistrationResult:
    """open3d ``registration_icp(source, target, th, init, PointToPoint(), criteria)`` on (n,3)
    arrays -- the operator behind cluster_icp.py:157, link.py:113 and evaluation.py:358."""
    if not max_correspondence_distance > 0:
        raise RuntimeError("[Open3D Error] Invalid max_correspondence_distance.")
    src = np.ascontiguousarray(source, dtype=np.float64).reshape(-1, 3)
    tgt = np.ascontiguousarray(target, dtype=np.float64).reshape(-1, 3)
    init = np.eye(4) if init is None else np.ascontiguousarray(init, dtype=np.float64).reshape(4, 4)
    r = _host_ctx().run(src, np.array([0, src.shape[0]], dtype=np.int32), tgt,
                        np.array([0, tgt.shape[0]], dtype=np.int32), np.zeros(1, dtype=np.int32), None, None,
                        init.reshape(1, 4, 4), box_scale=1.0, max_corr=max_correspondence_distance,
                        max_iter=max_iteration, rel_fitness=relative_fitness, rel_rmse=relative_rmse, want_world=False)
    corr = r["corr"]
    i = np.nonzero(corr >= 0)[0]
    return RegistrationResult(r["T"][0], float(r["fitness"][0]), float(r["rmse"][0]),
                              np.stack([i, corr[i]], 1).astype(np.int32), int(r["iters"][0]))
