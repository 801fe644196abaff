"""numpy restatement of /root/reference/PointCloud/dq_func.py (11 functions), of the four
pytorch3d 0.7.7 ``transforms.rotation_conversions`` functions it imports (dq_func.py:2),
of ``calculate_pc`` (mlp_reg.py:155-170) and of the local-frame move (mlp_reg.py:211-213).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for the pytorch3d pieces (library not vendored,
not installable; restated from its published algorithm and cross-checked against
scipy.spatial.transform.Rotation).  The dq_func layer itself IS pinned: tests/golden holds
outputs of the reference's own dq_func.py executed over these restated pytorch3d functions
(tests/golden/make_golden.py).  Arithmetic runs in the dtype of the input (float32 stays
float32), operation order as in the reference expressions.  Quaternions are real-first
(w, x, y, z), Hamilton convention.
"""
from __future__ import annotations

import numpy as np


# ---- pytorch3d.transforms.rotation_conversions (restated) -------------------------------

def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return np.stack((ow, ox, oy, oz), -1)


def quaternion_invert(q):
    return q * np.array([1, -1, -1, -1], dtype=q.dtype)


def quaternion_to_matrix(q):
    r, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two_s = q.dtype.type(2.0) / (q * q).sum(-1)
    o = np.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    ret = np.zeros_like(x)
    pos = x > 0
    ret[pos] = np.sqrt(x[pos])
    return ret


def matrix_to_quaternion(M):
    dt = M.dtype
    batch = M.shape[:-2]
    m = M.reshape(batch + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[..., i] for i in range(9)]
    one = dt.type(1.0)
    q_abs = _sqrt_positive_part(np.stack((one + m00 + m11 + m22, one + m00 - m11 - m22,
                                          one - m00 + m11 - m22, one - m00 - m11 + m22), -1))
    quat_by_rijk = np.stack((
        np.stack((q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01), -1),
        np.stack((m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20), -1),
        np.stack((m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21), -1),
        np.stack((m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2), -1)), -2)
    flr = dt.type(0.1)
    cand = quat_by_rijk / (dt.type(2.0) * np.maximum(q_abs[..., None], flr))
    best = np.argmax(q_abs, axis=-1)
    out = np.take_along_axis(cand, best[..., None, None], axis=-2)[..., 0, :]
    # standardize_quaternion: non-negative real part (present since pytorch3d 0.7.5)
    return np.where(out[..., 0:1] < 0, -out, out)


# ---- dq_func.py ------------------------------------------------------------------------

def transform_from_rot_trans(R, t):          # dq_func.py:4-27
    assert R.shape[-2:] == (3, 3)
    assert t.shape[-1] == 3
    T = np.zeros(R.shape[:-2] + (4, 4), dtype=R.dtype)
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T


def quaternion_conjugate(q):                 # :29-45
    assert q.shape[-1] == 4
    return np.concatenate([q[..., :1], -q[..., 1:]], -1)


def quat_trans_to_dualquat(q, t):            # :47-70
    assert q.shape[-1] == 4
    assert t.shape[-1] == 3
    q_dual = np.concatenate([np.zeros_like(q[..., :1]), t], -1)
    dq_dual = q.dtype.type(0.5) * quaternion_raw_multiply(q_dual, q)
    return np.concatenate([q, dq_dual], -1)


def rot_trans_to_dualquat(R, t):             # :72-98
    assert R.shape[-2:] == (3, 3)
    assert t.shape[-1] == 3
    eps = np.finfo(R.dtype).eps
    q_rot = matrix_to_quaternion(R)
    n = np.sqrt((q_rot * q_rot).sum(-1, keepdims=True))
    return quat_trans_to_dualquat(q_rot / np.maximum(n, eps), t)


def transform_to_dualquat(T):                # :100-124
    assert T.shape[-2:] == (4, 4)
    return rot_trans_to_dualquat(T[..., :3, :3], T[..., :3, 3])


def dualquat_to_quat_trans(dq):              # :126-146  (q = q_r * q_d, as written upstream)
    assert dq.shape[-1] == 8
    r, d = dq[..., :4], dq[..., 4:]
    q = quaternion_raw_multiply(r, d)
    t = dq.dtype.type(2) * quaternion_raw_multiply(d, quaternion_invert(r))
    return q, t[..., 1:]


def dualquat_to_rot_trans(dq):               # :148-168
    assert dq.shape[-1] == 8
    r, d = dq[..., :4], dq[..., 4:]
    R = quaternion_to_matrix(r)
    t = dq.dtype.type(2) * quaternion_raw_multiply(d, quaternion_invert(r))
    return R, t[..., 1:]


def dualquat_to_transform(dq):               # :170-186
    assert dq.shape[-1] == 8
    R, t = dualquat_to_rot_trans(dq)
    return transform_from_rot_trans(R, t)


def dualquat_multiply(a, b):                 # :188-211
    assert a.shape[-1] == 8
    assert b.shape[-1] == 8
    ar, ad, br, bd = a[..., :4], a[..., 4:], b[..., :4], b[..., 4:]
    return np.concatenate([quaternion_raw_multiply(ar, br),
                           quaternion_raw_multiply(ar, bd) + quaternion_raw_multiply(ad, br)], -1)


def dualquat_invert(dq):                     # :213-236
    assert dq.shape[-1] == 8
    eps = np.finfo(dq.dtype).eps
    r, d = dq[..., :4], dq[..., 4:]
    n2 = np.sqrt((r * r).sum(-1, keepdims=True)) ** 2
    rc = quaternion_conjugate(r)
    inv_r = rc / np.maximum(n2, eps)
    d_n = quaternion_conjugate(d) / np.maximum(n2, eps)
    dot = (r * d).sum(-1, keepdims=True) / np.maximum(n2, eps) ** 2
    return np.concatenate([inv_r, d_n - dq.dtype.type(2) * rc * dot], -1)


def point_to_dualquat(p):                    # :238-257
    assert p.shape[-1] == 3
    uq = np.zeros(p.shape[:-1] + (4,), dtype=p.dtype)
    uq[..., 0] = 1.0
    return np.concatenate([uq, np.zeros_like(p[..., :1]), p], -1)


# ---- SE(3) apply: mlp_reg.py:155-170 and :211-213 ----------------------------------------

def calculate_pc(local_clusters, matrices):
    return [ic @ matrices[i][:3, :3].T + matrices[i][:3, 3] for i, ic in enumerate(local_clusters)]


def to_local(points, matrix):
    """``inv(T) @ [X;1]`` rows 0..2 -- mlp_reg.py:211-213, cluster_icp.py:96-98 (float64)"""
    inv = np.linalg.inv(matrix)
    return (inv @ np.hstack([points, np.ones((points.shape[0], 1))]).T)[:3].T
