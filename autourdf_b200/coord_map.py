"""Pose bookkeeping between registration and structure inference: the part of reference
``CoordMap`` (PointCloud/coord_map.py:131-307) that reads what the registration loop wrote
(``matrix/{t:04}.npy``, ``cluster/{t:04}.npz``; mlp_reg.py:331-332) and turns it into the
pairwise cluster motion-distance map.  Same method names and return values; the O(T K^2)
Python loops of ``coord_dist_map`` (one roma / torch call per element upstream) run as three
CUDA kernels behind ``aurdf_coord_dist_map``.

Not mirrored (out of scope): the MST / silhouette clustering, the kinematic
tree, URDF and mesh generation, and ``get_bounding_box`` -- it needs open3d to read the raw
``.ply`` scans; pass the diagonal of their axis-aligned bounding box as ``bounding_box``.
"""
from __future__ import annotations

import glob

import numpy as np
import torch

from . import _lib, dq_func


def coord_dist_map(matrices, bounding_box: float, diff: bool = True):
    """coord_map.py:230-307.  ``matrices``: (T, K, 4, 4) float64 (numpy or CUDA tensor).
    Returns ``(coord_dist_map (K, K, T-1 if diff else T), sum_map (K, K))`` as float64 numpy
    arrays (CUDA tensors if the input was a CUDA tensor)."""
    L = _lib.lib()
    was_np = not torch.is_tensor(matrices)
    dev = torch.device("cuda", torch.cuda.current_device())
    M = torch.as_tensor(np.asarray(matrices, dtype=np.float64) if was_np else matrices).to(dev, torch.float64).contiguous()
    assert M.dim() == 4 and M.shape[-2:] == (4, 4), "matrices must be (time-step, num_coords, 4, 4)"
    T, K = int(M.shape[0]), int(M.shape[1])
    steps = max(T - 1, 0) if diff else T
    out_map = torch.empty((K, K, steps), dtype=torch.float64, device=dev)
    out_sum = torch.empty((K, K), dtype=torch.float64, device=dev)
    ws = torch.empty(int(L.aurdf_coord_dist_map_workspace_bytes(T, K, int(bool(diff)))), dtype=torch.uint8, device=dev)
    _lib.check(L.aurdf_coord_dist_map(_lib.ptr(M), T, K, float(bounding_box), int(bool(diff)), _lib.ptr(out_map),
                                      _lib.ptr(out_sum), _lib.ptr(ws), ws.numel(), _lib.current_stream()),
               "aurdf_coord_dist_map")
    if was_np:
        return out_map.cpu().numpy(), out_sum.cpu().numpy()
    return out_map, out_sum


class CoordMap:
    """Loader + motion-distance map of reference ``CoordMap`` (coord_map.py:131-307).

    - ``self.coords``:   (time-step, num_coords, 7) x, y, z + real-first quaternion
    - ``self.matrices``: (time-step, num_coords, 4, 4)
    - ``self.clusters``: list (time-step) of ``NpzFile`` keyed '0'..'K-1'
    """

    def __init__(self, data_path, bounding_box: float, start_steps=0, end_steps=0):
        self.data_path = data_path
        self.start_steps = start_steps
        self.end_steps = end_steps
        self.coords, self.matrices = self.load_matrix(start_steps, end_steps)
        self.clusters = self.load_cluster(start_steps, end_steps)
        self.num_coords = self.coords.shape[1]
        self.scale = self.get_scale()
        self.bounding_box = float(bounding_box)

    def get_scale(self):
        """coord_map.py:173-180"""
        return max(np.max(self.coords[0, :, i]) - np.min(self.coords[0, :, i]) for i in range(3))

    def load_matrix(self, start_steps=0, end_steps=0):
        """coord_map.py:186-220: every pose as xyz + quaternion (one batched CUDA conversion instead
        of a torch call per matrix) and the stacked matrices."""
        files = sorted(glob.glob(self.data_path + "matrix/*.npy"))
        files = files[start_steps:end_steps]
        matrices = np.array([np.load(f) for f in files])
        if matrices.size == 0:          # upstream: the loops do not run, both results are empty arrays
            return np.array([]), matrices
        m = torch.as_tensor(matrices).cuda()
        q = dq_func.matrix_to_quaternion(m[..., :3, :3].contiguous())
        rot = torch.cat([m[..., :3, 3], q], dim=-1).cpu().numpy()
        return rot, matrices

    def load_cluster(self, start_steps=0, end_steps=0):
        """coord_map.py:221-228"""
        files = sorted(glob.glob(self.data_path + "cluster/*.npz"))
        files = files[start_steps:end_steps]
        return [np.load(f) for f in files]

    def coord_dist_map(self, diff=True):
        """coord_map.py:230-307: (num_seg, num_seg, time-step) map and its sum over time"""
        return coord_dist_map(self.matrices, self.bounding_box, diff)
