"""SE(3) apply operators of the registration loop (reference PointCloud/mlp_reg.py).

``calculate_pc`` keeps the reference signature (mlp_reg.py:155-170) and stays
autograd-compatible -- ``train()`` differentiates through it (mlp_reg.py:93-116) -- by
wrapping the batched CUDA kernels ``aurdf_se3_apply`` / ``aurdf_se3_apply_bwd`` in a
``torch.autograd.Function``: all K clusters go through ONE launch instead of K small
matmuls.  ``to_local`` is the inverse move of mlp_reg.py:211-213 / cluster_icp.py:96-98.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_TORCH2DT = {torch.float32: _lib.F32, torch.float64: _lib.F64}


class _Se3Apply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, off, T):
        L = _lib.lib()
        xyz = xyz.contiguous()
        T = T.contiguous()
        out = torch.empty_like(xyz)
        _lib.check(L.aurdf_se3_apply(_lib.ptr(xyz), _lib.ptr(off), _lib.ptr(T), T.shape[0], xyz.shape[0],
                                     _TORCH2DT[xyz.dtype], _lib.ptr(out), _lib.current_stream()), "aurdf_se3_apply")
        ctx.save_for_backward(xyz, off, T)
        return out

    @staticmethod
    def backward(ctx, g):
        xyz, off, T = ctx.saved_tensors
        L = _lib.lib()
        g = g.contiguous()
        gx = torch.empty_like(xyz) if ctx.needs_input_grad[0] else None
        gT = torch.empty_like(T) if ctx.needs_input_grad[2] else None
        _lib.check(L.aurdf_se3_apply_bwd(_lib.ptr(g), _lib.ptr(xyz), _lib.ptr(off), _lib.ptr(T), T.shape[0],
                                         xyz.shape[0], _TORCH2DT[xyz.dtype], _lib.ptr(gx), _lib.ptr(gT),
                                         _lib.current_stream()), "aurdf_se3_apply_bwd")
        return gx, None, gT


def se3_apply(xyz: torch.Tensor, off: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """packed form: xyz (N,3), off (K+1,) int32, T (K,4,4) -> (N,3); differentiable in xyz and T"""
    assert xyz.is_cuda and T.is_cuda and off.is_cuda and off.dtype == torch.int32
    assert xyz.dtype == T.dtype and xyz.dtype in _TORCH2DT and T.shape[-2:] == (4, 4)
    return _Se3Apply.apply(xyz, off, T)


def _offsets(sizes, device):
    off = np.zeros(len(sizes) + 1, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    return torch.from_numpy(off).to(device)


def calculate_pc(local_clusters, matrices):
    """Reference signature (mlp_reg.py:155): list of (n_k,3) tensors + (K,4,4) tensor ->
    list of (n_k,3) tensors in the world frame."""
    K = len(local_clusters)
    if K == 0:
        return []
    sizes = [int(c.shape[0]) for c in local_clusters]
    xyz = torch.cat(local_clusters, dim=0)
    out = se3_apply(xyz, _offsets(sizes, xyz.device), matrices[:K])
    return list(torch.split(out, sizes, dim=0))


def to_local(points_list, matrices):
    """``(inv(T_k) @ [X_k;1])[:3].T`` for every cluster -- mlp_reg.py:211-213.  numpy float64
    in / out (as upstream); computed by ``aurdf_se3_to_local``."""
    L = _lib.lib()
    K = len(points_list)
    sizes = [int(np.asarray(p).reshape(-1, 3).shape[0]) for p in points_list]
    if K == 0 or sum(sizes) == 0:
        return [np.zeros((0, 3)) for _ in range(K)]
    dev = torch.device("cuda")
    xyz = torch.from_numpy(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in points_list])).to(dev)
    T = torch.from_numpy(np.ascontiguousarray(np.asarray(matrices, dtype=np.float64)[:K].reshape(K, 4, 4))).to(dev)
    out = torch.empty_like(xyz)
    off = _offsets(sizes, dev)
    _lib.check(L.aurdf_se3_to_local(_lib.ptr(xyz), _lib.ptr(off), _lib.ptr(T), K, xyz.shape[0], _lib.ptr(out),
                                    _lib.current_stream()), "aurdf_se3_to_local")
    o = out.cpu().numpy()
    cuts = np.cumsum(sizes)[:-1]
    return [a.copy() for a in np.split(o, cuts)]


def resample_cluster(segments, idx, n_clusters, matrices, normal=False, visual=False, _details=None):
    """Reference signature (mlp_reg.py:172): ``segments`` is the reference's ``Segments`` object
    (its ``pc_list[idx].points`` is the frame cloud) or directly an (N,3) array.  Seeded Lloyd
    k-means + move into local frames, all on the GPU (``aurdf_resample_clusters``).  Returns the
    list of (n_k,3) float64 clusters in their local frames.  ``normal=True`` needs open3d normal
    estimation and is not on this path; ``visual`` is accepted and ignored."""
    if normal:
        raise NotImplementedError("normal=True (open3d normal estimation) is outside the accelerated path")
    if hasattr(segments, "pc_list"):
        pc_np = np.asarray(segments.pc_list[idx].points)
    else:
        pc_np = np.asarray(segments)
    pc_np = np.ascontiguousarray(pc_np, dtype=np.float64).reshape(-1, 3)
    mats = np.ascontiguousarray(np.asarray(matrices, dtype=np.float64)[:n_clusters].reshape(n_clusters, 4, 4))
    L = _lib.lib()
    dev = torch.device("cuda")
    N = pc_np.shape[0]
    cloud = torch.from_numpy(pc_np).to(dev)
    off = torch.tensor([0, N], dtype=torch.int32, device=dev)
    labels = torch.empty(N, dtype=torch.int32, device=dev)
    centers = torch.empty((n_clusters, 3), dtype=torch.float64, device=dev)
    local = torch.empty((N, 3), dtype=torch.float64, device=dev)
    loff = torch.empty(n_clusters + 1, dtype=torch.int32, device=dev)
    nit = torch.empty(1, dtype=torch.int32, device=dev)
    inertia = torch.empty(1, dtype=torch.float64, device=dev)
    _lib.check(L.aurdf_resample_clusters(_lib.ptr(cloud), _lib.ptr(off), _lib.ptr(torch.from_numpy(mats).to(dev)), 1,
                                         n_clusters, N, 300, 1e-4, _lib.ptr(labels), _lib.ptr(centers), _lib.ptr(local),
                                         _lib.ptr(loff), _lib.ptr(nit), _lib.ptr(inertia), _lib.current_stream()),
               "aurdf_resample_clusters")
    lo = loff.cpu().numpy()
    loc = local.cpu().numpy()
    if _details is not None:
        _details.update(labels=labels.cpu().numpy(), centers=centers.cpu().numpy(), n_iter=int(nit.item()),
                        inertia=float(inertia.item()))
    return [loc[lo[k]:lo[k + 1]].copy() for k in range(n_clusters)]
