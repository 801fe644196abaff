"""pytorch3d 0.7.7 ``transforms.rotation_conversions`` (the four functions the reference imports at
PointCloud/dq_func.py:2 and mlp_reg.py:13) and ``dq_func.py``'s live pair, restated in plain torch.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: pytorch3d is not installable here, so these follow the
published 0.7.7 algorithm; they are differentiable through torch autograd, which is what makes them the
reference for the CUDA backward kernels (``aurdf_dq_op_bwd``) and the stand-in the golden generators put
behind the reference's own ``train()`` (tests/golden/make_golden*.py)."""
from __future__ import annotations

import torch


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def quaternion_invert(q):
    return q * q.new_tensor([1, -1, -1, -1])


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(matrix):
    batch = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
                                             1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    cand = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    idx = q_abs.argmax(dim=-1)
    out = torch.gather(cand, -2, idx[..., None, None].expand(batch + (1, 4)))[..., 0, :]
    return torch.where(out[..., 0:1] < 0, -out, out)


def transform_to_dualquat(T):
    """dq_func.py:100-124 -> :72-98 -> :47-70"""
    R, t = T[..., :3, :3], T[..., :3, 3]
    q = matrix_to_quaternion(R)
    q = q / torch.clamp(torch.norm(q, dim=-1, keepdim=True), min=torch.finfo(q.dtype).eps)
    qd = torch.cat([torch.zeros_like(t[..., :1]), t], dim=-1)
    return torch.cat([q, 0.5 * quaternion_raw_multiply(qd, q)], dim=-1)


def dualquat_to_transform(dq):
    """dq_func.py:170-186 -> :148-168"""
    qr, qd = dq[..., :4], dq[..., 4:]
    R = quaternion_to_matrix(qr)
    t = 2.0 * quaternion_raw_multiply(qd, quaternion_invert(qr))[..., 1:]
    T = torch.zeros(dq.shape[:-1] + (4, 4), dtype=dq.dtype, device=dq.device)
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1.0
    return T
