"""CPU restatement of the pose bookkeeping that consumes the registration output
(SURVEY.md section 8(f)-4): the on-disk formats and the motion-distance map.

TEST INFRASTRUCTURE ONLY (tests/, bench scripts' CPU legs).  Follows

  * /root/reference/PointCloud/helper_functions.py:10-21   save_pc_npz / load_pc_npz
  * /root/reference/PointCloud/coord_map.py:186-220        CoordMap.load_matrix
  * /root/reference/PointCloud/coord_map.py:230-307        CoordMap.coord_dist_map

PARITY UNPINNED at the third-party boundary: coord_dist_map calls ``roma.rotmat_to_rotvec``,
``roma.utils.rotvec_geodesic_distance`` and ``roma.rotmat_geodesic_distance``; roma is imported
by the reference (coord_map.py:14) but neither listed in its requirements.txt nor installed
here, so its published algorithms are restated below (they are scipy's ``Rotation`` algorithms,
which roma says it adapted, so scipy pins them independently in tests/test_coord_map.py):

  rotmat_to_unitquat   scipy ``from_matrix -> as_quat`` branch selection on (diag, trace), xyzw
  unitquat_to_rotvec   shortest arc (w >= 0), angle = 2 atan2(|xyz|, w), series below 1e-3
  rotvec_to_unitquat   scipy ``from_rotvec`` (series below 1e-3)
  unitquat_geodesic_distance(q1, q2) = 4 asin(min(|q2 - q1|, |q2 + q1|) / 2)
  rotmat_geodesic_distance(R1, R2)   = 2 asin(min(|R2 - R1|_F / (2 sqrt 2), 1))

The reference's own loops (which array is overwritten when, the three lambdas, the row-wise
L2 norms) are pinned by tests/golden/coord_map.npz, produced by the reference's
``CoordMap.coord_dist_map`` itself (tests/golden/make_golden_coord_map.py).
"""
from __future__ import annotations

import glob
import math

import numpy as np


# ---------------------------------------------------------------- roma, restated (float64)
def rotmat_to_unitquat(R):
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    n = R.shape[0]
    dec = np.empty((n, 4))
    dec[:, :3] = np.diagonal(R, axis1=1, axis2=2)
    dec[:, 3] = dec[:, :3].sum(1)
    choice = dec.argmax(1)
    q = np.empty((n, 4))
    ind = np.nonzero(choice != 3)[0]
    i = choice[ind]
    j = (i + 1) % 3
    k = (j + 1) % 3
    q[ind, i] = 1 - dec[ind, 3] + 2 * R[ind, i, i]
    q[ind, j] = R[ind, j, i] + R[ind, i, j]
    q[ind, k] = R[ind, k, i] + R[ind, i, k]
    q[ind, 3] = R[ind, k, j] - R[ind, j, k]
    ind = np.nonzero(choice == 3)[0]
    q[ind, 0] = R[ind, 2, 1] - R[ind, 1, 2]
    q[ind, 1] = R[ind, 0, 2] - R[ind, 2, 0]
    q[ind, 2] = R[ind, 1, 0] - R[ind, 0, 1]
    q[ind, 3] = 1 + dec[ind, 3]
    return q / np.linalg.norm(q, axis=1)[:, None]


def unitquat_to_rotvec(q):
    q = np.array(q, dtype=np.float64).reshape(-1, 4)
    q[q[:, 3] < 0] *= -1
    half = np.arctan2(np.linalg.norm(q[:, :3], axis=1), q[:, 3])
    ang = 2 * half
    small = np.abs(ang) <= 1e-3
    scale = np.empty(q.shape[0])
    scale[small] = 2 + ang[small] ** 2 / 12 + 7 * ang[small] ** 4 / 2880
    scale[~small] = ang[~small] / np.sin(half[~small])
    return scale[:, None] * q[:, :3]


def rotmat_to_rotvec(R):
    return unitquat_to_rotvec(rotmat_to_unitquat(R))


def rotvec_to_unitquat(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
    ang = np.linalg.norm(v, axis=1)
    small = ang <= 1e-3
    scale = np.empty(v.shape[0])
    scale[small] = 0.5 - ang[small] ** 2 / 48 + ang[small] ** 4 / 3840
    scale[~small] = np.sin(ang[~small] / 2) / ang[~small]
    return np.concatenate([scale[:, None] * v, np.cos(ang / 2)[:, None]], 1)


def unitquat_geodesic_distance(q1, q2):
    return 4.0 * np.arcsin(0.5 * np.minimum(np.linalg.norm(q2 - q1, axis=-1), np.linalg.norm(q2 + q1, axis=-1)))


def rotvec_geodesic_distance(v1, v2):
    return unitquat_geodesic_distance(rotvec_to_unitquat(v1), rotvec_to_unitquat(v2))


def rotmat_geodesic_distance(R1, R2):
    d = np.linalg.norm(np.asarray(R2) - np.asarray(R1), axis=(-2, -1)) / (2.0 * math.sqrt(2.0))
    return 2.0 * np.arcsin(np.minimum(d, 1.0))


# ---------------------------------------------------------------- pytorch3d matrix_to_quaternion (real first)
def matrix_to_quaternion(M):
    from .dq_oracle import matrix_to_quaternion as m2q
    return m2q(np.asarray(M, dtype=np.float64))


# ---------------------------------------------------------------- file formats
def save_pc_npz(segment_list, path):
    """helper_functions.py:10-16 -- keys are the decimal strings '0'..'K-1'"""
    np.savez(path, **{f"{i}": pc for i, pc in enumerate(segment_list)})


def load_pc_npz(path):
    """helper_functions.py:18-21"""
    z = np.load(path)
    return [z[k] for k in z.keys()]


def load_matrix(data_path, start_steps=0, end_steps=0):
    """coord_map.py:186-220: (T,K,7) xyz + real-first quaternion, and the (T,K,4,4) matrices"""
    files = sorted(glob.glob(data_path + "matrix/*.npy"))[start_steps:end_steps]
    matrices = np.array([np.load(f) for f in files])
    if matrices.size == 0:          # upstream: the loops do not run, both results are empty arrays
        return np.array([]), matrices
    T, K = matrices.shape[:2]
    q = matrix_to_quaternion(matrices[:, :, :3, :3].reshape(-1, 3, 3)).reshape(T, K, 4)
    return np.concatenate([matrices[:, :, :3, 3], q], -1), matrices


# ---------------------------------------------------------------- motion-distance map
def coord_dist_map(matrices, bounding_box, diff=True):
    """coord_map.py:230-307.  matrices (T,K,4,4) float64 -> (K,K,T-1 or T), (K,K)"""
    M = np.asarray(matrices, dtype=np.float64)
    T, K = M.shape[:2]
    lam_rot = 1 / math.pi
    lam_bbox = 1 / (bounding_box * 2)
    xyz = M[:, :, :3, 3]
    maps = []
    if diff:
        trans_diff = np.diff(xyz, axis=0)
        for i in range(T - 1):
            rel = np.einsum("kji,kjl->kil", M[i, :, :3, :3], M[i + 1, :, :3, :3])      # R_i^T R_{i+1}
            rv = rotmat_to_rotvec(rel)
            d_xyz = lam_bbox * np.linalg.norm(trans_diff[i][:, None, :] - trans_diff[i][None, :, :], axis=-1)
            d_rpy = lam_rot * rotvec_geodesic_distance(np.repeat(rv, K, 0), np.tile(rv, (K, 1))).reshape(K, K)
            trans_dist = np.linalg.norm(d_xyz[:, None, :] - d_xyz[None, :, :], axis=-1)
            rot_dist = np.linalg.norm(d_rpy[:, None, :] - d_rpy[None, :, :], axis=-1)
            maps.append(trans_dist + rot_dist)
    else:
        for i in range(T):
            d_xyz = lam_bbox * np.linalg.norm(xyz[i][:, None, :] - xyz[i][None, :, :], axis=-1)
            R = M[i, :, :3, :3]
            d_rpy = lam_rot * rotmat_geodesic_distance(R[:, None], R[None, :])
            maps.append(d_xyz + d_rpy)
    cmap = np.stack(maps, axis=2) if maps else np.zeros((K, K, 0))
    return cmap, np.sum(np.abs(cmap), axis=2)
