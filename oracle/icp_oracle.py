"""ctypes front-end of ``oracle/icp_oracle.c`` plus the reference-shaped entry points.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED at the open3d
boundary; the outer sweep follows /root/reference/PointCloud/cluster_icp.py:118-191.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile ``libicp_oracle.so`` in place (gcc, a second or two)."""
    so = os.path.join(_HERE, "libicp_oracle.so")
    src = os.path.join(_HERE, "icp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libicp_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_svd3.argtypes = [_dp, _dp, _dp, _dp]
        L.orc_kabsch.argtypes = [_dp, C.c_int, _dp, _ip, _dp]
        L.orc_nn_batch.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, _ip, _dp]
        L.orc_icp_p2p.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, _dp, C.c_int, C.c_double, C.c_double,
                                  C.c_int, _dp, _ip, _dp, _dp, _ip, _dp]
        L.orc_icp_p2p.restype = C.c_int
        L.orc_aabb_box.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _dp]
        L.orc_mask.argtypes = [_dp, C.c_int, _dp, _ip]
        L.orc_mask.restype = C.c_int
        L.orc_transform_pts.argtypes = [_dp, _dp, C.c_int, _dp]
        L.orc_mat4_mul.argtypes = [_dp, _dp, _dp]
        L.orc_mat4_inv.argtypes = [_dp, _dp]
        L.orc_mat4_inv.restype = C.c_int
        L.orc_masked_icp_sweep.argtypes = [_dp, _ip, _dp, _ip, _ip, C.c_void_p, C.c_int, _ip, _dp, C.c_int,
                                           C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                           C.c_int, C.c_int, _dp, _dp, _ip, _dp, _dp, _ip, _ip, _dp]
        L.orc_masked_icp_sweep.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_set_reference_threading.argtypes = [C.c_int]
        L.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def set_reference_threading(on: bool):
    """True: tiles serial, OpenMP inside the correspondence search (open3d's own structure);
    False (default): OpenMP over tiles, deterministic per-tile arithmetic."""
    lib().orc_set_reference_threading(int(bool(on)))


def use_all_host_threads() -> int:
    """OpenMP thread count := CPUs this process may run on (torchrun exports OMP_NUM_THREADS=1)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(n)
    return n


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None:
        a = a.reshape(-1, shape_last)
    return a


def svd3(A):
    A = _f64(A).reshape(3, 3)
    U = np.empty((3, 3)); S = np.empty(3); V = np.empty((3, 3))
    lib().orc_svd3(_d(A), _d(U), _d(S), _d(V))
    return U, S, V


def kabsch(P, Q, corr):
    P = _f64(P, 3); Q = _f64(Q, 3)
    corr = np.ascontiguousarray(corr, dtype=np.int32)
    U = np.empty((4, 4))
    lib().orc_kabsch(_d(P), P.shape[0], _d(Q), _i(corr), _d(U))
    return U


def nn_batch(P, Q, use_kdtree=False):
    P = _f64(P, 3); Q = _f64(Q, 3)
    idx = np.empty(P.shape[0], dtype=np.int32); d2 = np.empty(P.shape[0])
    lib().orc_nn_batch(_d(P), P.shape[0], _d(Q), Q.shape[0], int(use_kdtree), _i(idx), _d(d2))
    return idx, d2


def transform_pts(T, pts):
    T = _f64(T).reshape(4, 4); pts = _f64(pts, 3)
    out = np.empty_like(pts)
    lib().orc_transform_pts(_d(T), _d(pts), pts.shape[0], _d(out))
    return out


def mat4_inv(T):
    T = _f64(T).reshape(4, 4)
    out = np.empty((4, 4))
    if lib().orc_mat4_inv(_d(T), _d(out)) != 0:
        raise np.linalg.LinAlgError("Singular matrix")
    return out


def icp_p2p(src, tgt, max_corr, init, max_iter=30, rel_fit=1e-6, rel_rmse=1e-6, use_kdtree=False):
    """open3d ``registration_icp(src, tgt, max_corr, init, PointToPoint(), criteria)`` restated.

    Returns dict(T, corr, fitness, rmse, iters, P).  ``max_iter`` default 30 is open3d's;
    the reference passes 10000 (cluster_icp.py:159)."""
    src = _f64(src, 3); tgt = _f64(tgt, 3); T0 = _f64(init).reshape(4, 4)
    ns = src.shape[0]
    T = np.empty((4, 4)); corr = np.empty(max(ns, 1), dtype=np.int32); P = np.empty((max(ns, 1), 3))
    fit = C.c_double(); rmse = C.c_double(); iters = C.c_int()
    rc = lib().orc_icp_p2p(_d(src), ns, _d(tgt), tgt.shape[0], float(max_corr), _d(T0), int(max_iter),
                           float(rel_fit), float(rel_rmse), int(use_kdtree), _d(T), _i(corr), C.byref(fit),
                           C.byref(rmse), C.byref(iters), _d(P))
    if rc == -1:
        raise RuntimeError("[Open3D Error] Invalid max_correspondence_distance.")
    if rc != 0:
        raise MemoryError
    return dict(T=T, corr=corr[:ns], fitness=fit.value, rmse=rmse.value, iters=iters.value, P=P[:ns])


def aabb_box(pts, scale=1.2):
    """cluster_icp.py:133-140 -> (lo[3], hi[3]) as float64; float32 arithmetic for float32 input."""
    pts = np.asarray(pts)
    is_f32 = pts.dtype == np.float32
    pts = np.ascontiguousarray(pts, dtype=np.float32 if is_f32 else np.float64).reshape(-1, 3)
    box = np.empty(6)
    lib().orc_aabb_box(pts.ctypes.data_as(C.c_void_p), pts.shape[0], int(is_f32), float(scale), _d(box))
    return box[:3].copy(), box[3:].copy()


def mask(cloud, lo, hi):
    cloud = _f64(cloud, 3)
    box = np.concatenate([lo, hi]).astype(np.float64)
    idx = np.empty(max(cloud.shape[0], 1), dtype=np.int32)
    n = lib().orc_mask(_d(cloud), cloud.shape[0], _d(box), _i(idx))
    return idx[:n].copy()


def pack(arrs, dtype):
    """list of (n_k,3) arrays -> packed (sum n,3) array + int32 offsets[K+1]"""
    off = np.zeros(len(arrs) + 1, dtype=np.int32)
    for k, a in enumerate(arrs):
        off[k + 1] = off[k] + np.asarray(a).reshape(-1, 3).shape[0]
    out = np.empty((int(off[-1]), 3), dtype=dtype)
    for k, a in enumerate(arrs):
        out[off[k]:off[k + 1]] = np.asarray(a).reshape(-1, 3)
    return out, off


def masked_icp_sweep(src, src_off, tgt, tgt_off, tile_frame, box_pts, box_off, init_T, box_scale=1.2,
                     max_corr=1.0, max_iter=10000, rel_fit=1e-6, rel_rmse=1e-6, ori_only=False,
                     use_kdtree=False, nthreads=0):
    """Batched sweep over (frame, cluster) tiles -- the packed form of ``masked_icp``."""
    src = _f64(src, 3); tgt = _f64(tgt, 3)
    src_off = np.ascontiguousarray(src_off, dtype=np.int32)
    tgt_off = np.ascontiguousarray(tgt_off, dtype=np.int32)
    tile_frame = np.ascontiguousarray(tile_frame, dtype=np.int32)
    box_off = np.ascontiguousarray(box_off, dtype=np.int32)
    box_pts = np.asarray(box_pts)
    is_f32 = box_pts.dtype == np.float32
    box_pts = np.ascontiguousarray(box_pts, dtype=np.float32 if is_f32 else np.float64).reshape(-1, 3)
    init_T = _f64(init_T).reshape(-1, 16)
    B = tile_frame.shape[0]
    n = src.shape[0]
    out_T = np.empty((B, 4, 4)); out_world = np.empty((max(n, 1), 3)); out_corr = np.empty(max(n, 1), dtype=np.int32)
    out_fit = np.empty(max(B, 1)); out_rmse = np.empty(max(B, 1))
    out_iters = np.empty(max(B, 1), dtype=np.int32); out_ntgt = np.empty(max(B, 1), dtype=np.int32)
    out_cond = np.ones(max(B, 1))
    rc = lib().orc_masked_icp_sweep(_d(src), _i(src_off), _d(tgt), _i(tgt_off), _i(tile_frame),
                                    box_pts.ctypes.data_as(C.c_void_p), int(is_f32), _i(box_off), _d(init_T), B,
                                    float(box_scale), float(max_corr), int(max_iter), float(rel_fit),
                                    float(rel_rmse), int(bool(ori_only)), int(use_kdtree), int(nthreads),
                                    _d(out_T), _d(out_world), _i(out_corr), _d(out_fit), _d(out_rmse),
                                    _i(out_iters), _i(out_ntgt), _d(out_cond))
    if rc == -1:
        raise RuntimeError("[Open3D Error] Invalid max_correspondence_distance.")
    if rc != 0:
        raise MemoryError
    return dict(T=out_T, world=out_world[:n], corr=out_corr[:n], fitness=out_fit[:B], rmse=out_rmse[:B],
                iters=out_iters[:B], ntgt=out_ntgt[:B], cond=out_cond[:B])


def masked_icp(clusters_local, clusters_world, step_pc_np, matrices, visual=False, ori=False, scale=1.2, th=1,
               colors=None, _details=None, use_kdtree=False):
    """Reference signature (cluster_icp.py:118).  Returns (list of (n_k,3) f64, (K,4,4) f64)."""
    K = min(len(clusters_local), len(clusters_world), len(matrices))  # zip() truncation, :131
    clusters_local = [np.asarray(c) for c in clusters_local[:K]]
    clusters_world = [np.asarray(c) for c in clusters_world[:K]]
    for c in clusters_world:
        if c.reshape(-1, 3).shape[0] == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
    wdt = np.float32 if all(c.dtype == np.float32 for c in clusters_world) and K > 0 else np.float64
    src, src_off = pack(clusters_local, np.float64)
    box, box_off = pack(clusters_world, wdt)
    tgt = _f64(step_pc_np, 3)
    tgt_off = np.array([0, tgt.shape[0]], dtype=np.int32)
    init_T = np.asarray([np.asarray(m, dtype=np.float64) for m in matrices[:K]]).reshape(K, 16) if K else np.zeros((0, 16))
    r = masked_icp_sweep(src, src_off, tgt, tgt_off, np.zeros(K, dtype=np.int32), box, box_off, init_T,
                         box_scale=scale, max_corr=th, max_iter=10000, ori_only=ori, use_kdtree=use_kdtree)
    if _details is not None:
        _details.update(r, src_off=src_off)
    world = [r["world"][src_off[k]:src_off[k + 1]].copy() for k in range(K)]
    return world, r["T"].copy()
