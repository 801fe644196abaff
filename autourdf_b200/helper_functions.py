"""On-disk formats and pose-vector helpers of the registration output: reference
PointCloud/helper_functions.py:10-45, same names and behaviour.

``save_pc_npz`` / ``load_pc_npz`` are the ``cluster/{t:04}.npz`` format (keys are the decimal
strings '0'..'K-1', loaded back in stored order); they are plain numpy, as upstream.  The two
matrix <-> (xyz, quaternion) helpers call the CUDA ``matrix_to_quaternion`` /
``quaternion_to_matrix`` operators of :mod:`autourdf_b200.dq_func` (real-first quaternions, as
pytorch3d returns them -- the upstream docstrings say qx,qy,qz,qw but the code is w-first).
"""
from __future__ import annotations

import numpy as np
import torch

from . import dq_func


def save_pc_npz(segment_list, path):
    """helper_functions.py:10-16"""
    np.savez(path, **{f"{i}": pc_np for i, pc_np in enumerate(segment_list)})


def load_pc_npz(path):
    """helper_functions.py:18-21"""
    pc_npz = np.load(path)
    return [pc_npz[key] for key in pc_npz.keys()]


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def matrix2xyzquant_torch(matrix):
    """helper_functions.py:26-34: 4x4 transform -> 7-vector (x, y, z, qw, qx, qy, qz)"""
    m = torch.as_tensor(matrix)
    q = dq_func.matrix_to_quaternion(m[:3, :3].to(_dev()).contiguous()).to(m.device).squeeze()
    return torch.cat([m[:3, 3], q])


def xyzquant2matrix_torch(xyzquat):
    """helper_functions.py:36-45: 7-vector -> 4x4 transform (float32 ``torch.eye`` container, as upstream)"""
    v = torch.as_tensor(xyzquat)
    rot = dq_func.quaternion_to_matrix(v[3:].unsqueeze(0).to(_dev()).contiguous()).to(v.device).squeeze()
    matrix = torch.eye(4)
    matrix[:3, 3] = v[:3]
    matrix[:3, :3] = rot
    return matrix
