"""Dual-quaternion SE(3) library: the 11 functions of reference PointCloud/dq_func.py:4-257
with the same names, argument order, broadcasting batch dims and shape asserts, computed by
one batched CUDA kernel (``aurdf_dq_op``).  Also the four pytorch3d 0.7.7 functions the
reference imports (dq_func.py:2).  Real-first quaternions (w, x, y, z).

Differentiable: the reference backpropagates through ``matrix_to_quaternion`` /
``quaternion_to_matrix`` (``--r q``, mlp_reg.py:60-66) and ``transform_to_dualquat`` /
``dualquat_to_transform`` (``--r dq``, :78-84) between the pose MLP and ``loss.backward()``
(:114-116).  Every operator here is a ``torch.autograd.Function`` whose backward is the CUDA
vector-Jacobian kernel ``aurdf_dq_op_bwd`` (exact derivative of the same expression, on the branch
the forward pass took -- the subgradient torch autograd uses for the reference's code).
"""
from __future__ import annotations

import torch

from . import _lib

_TORCH2DT = {torch.float32: _lib.F32, torch.float64: _lib.F64}
(_OP_T_FROM_RT, _OP_QCONJ, _OP_QT2DQ, _OP_RT2DQ, _OP_T2DQ, _OP_DQ2QT, _OP_DQ2RT, _OP_DQ2T, _OP_DQMUL, _OP_DQINV,
 _OP_P2DQ, _OP_QMUL, _OP_QINV, _OP_Q2M, _OP_M2Q) = range(15)


class _DqOp(torch.autograd.Function):
    """one batched operator of ``aurdf_dq_op``; inputs are contiguous and already broadcast"""

    @staticmethod
    def forward(ctx, op, n, out_shape, out1_shape, in0, in1):
        L = _lib.lib()
        out0 = torch.empty(out_shape, dtype=in0.dtype, device=in0.device)
        out1 = torch.empty(out1_shape, dtype=in0.dtype, device=in0.device) if out1_shape is not None else None
        _lib.check(L.aurdf_dq_op(op, _lib.ptr(in0), _lib.ptr(in1), _lib.ptr(out0), _lib.ptr(out1), n,
                                 _TORCH2DT[in0.dtype], _lib.current_stream()), "aurdf_dq_op")
        ctx.op, ctx.n = op, n
        ctx.save_for_backward(in0, in1)
        if out1 is None:
            return out0
        return out0, out1

    @staticmethod
    def backward(ctx, g0, g1=None):
        L = _lib.lib()
        in0, in1 = ctx.saved_tensors
        need0, need1 = ctx.needs_input_grad[4], in1 is not None and ctx.needs_input_grad[5]
        if not (need0 or need1):
            return None, None, None, None, None, None
        g0 = g0.contiguous() if g0 is not None else None
        g1 = g1.contiguous() if g1 is not None else None
        gin0 = torch.empty_like(in0) if need0 else None
        gin1 = torch.empty_like(in1) if need1 else None
        _lib.check(L.aurdf_dq_op_bwd(ctx.op, _lib.ptr(in0), _lib.ptr(in1), _lib.ptr(g0), _lib.ptr(g1), _lib.ptr(gin0),
                                     _lib.ptr(gin1), ctx.n, _TORCH2DT[in0.dtype], _lib.current_stream()), "aurdf_dq_op_bwd")
        return None, None, None, None, gin0, gin1


def _run(op, in0, in_tail, in1, in1_tail, out_tail, out1_tail=None):
    """flatten batch dims, launch, reshape; *_tail = trailing element shape of each operand.
    Broadcasting / dtype conversion happen in torch (differentiable), the operator itself in CUDA."""
    assert in0.is_cuda and in0.dtype in _TORCH2DT, "dq_func operators need CUDA float32/float64 tensors"
    batch = in0.shape[:len(in0.shape) - len(in_tail)]
    if in1 is not None:
        b1 = in1.shape[:len(in1.shape) - len(in1_tail)]
        if b1 != batch:
            batch = torch.broadcast_shapes(batch, b1)
            in1 = in1.expand(batch + tuple(in1_tail))
            in0 = in0.expand(batch + tuple(in_tail))
        in1 = in1.to(in0.dtype).contiguous()
    in0 = in0.contiguous()
    n = 1
    for d in batch:
        n *= int(d)
    return _DqOp.apply(op, n, tuple(batch) + tuple(out_tail), tuple(batch) + tuple(out1_tail) if out1_tail else None,
                       in0, in1)


# ---- pytorch3d.transforms (dq_func.py:2) -------------------------------------------------
def quaternion_raw_multiply(a, b):
    return _run(_OP_QMUL, a, (4,), b, (4,), (4,))


def quaternion_invert(q):
    return _run(_OP_QINV, q, (4,), None, None, (4,))


def quaternion_to_matrix(q):
    return _run(_OP_Q2M, q, (4,), None, None, (3, 3))


def matrix_to_quaternion(M):
    return _run(_OP_M2Q, M, (3, 3), None, None, (4,))


# ---- dq_func.py --------------------------------------------------------------------------
def transform_from_rot_trans(R: torch.Tensor, t: torch.Tensor):
    assert R.shape[-2:] == (3, 3)
    assert t.shape[-1] == 3
    return _run(_OP_T_FROM_RT, R, (3, 3), t, (3,), (4, 4))


def quaternion_conjugate(q: torch.Tensor) -> torch.Tensor:
    assert q.shape[-1] == 4
    return _run(_OP_QCONJ, q, (4,), None, None, (4,))


def quat_trans_to_dualquat(q: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    assert q.shape[-1] == 4
    assert t.shape[-1] == 3
    return _run(_OP_QT2DQ, q, (4,), t, (3,), (8,))


def rot_trans_to_dualquat(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    assert R.shape[-2:] == (3, 3)
    assert t.shape[-1] == 3
    return _run(_OP_RT2DQ, R, (3, 3), t, (3,), (8,))


def transform_to_dualquat(T: torch.Tensor) -> torch.Tensor:
    assert T.shape[-2:] == (4, 4)
    return _run(_OP_T2DQ, T, (4, 4), None, None, (8,))


def dualquat_to_quat_trans(dq: torch.Tensor):
    assert dq.shape[-1] == 8
    return _run(_OP_DQ2QT, dq, (8,), None, None, (4,), (3,))


def dualquat_to_rot_trans(dq: torch.Tensor):
    assert dq.shape[-1] == 8
    return _run(_OP_DQ2RT, dq, (8,), None, None, (3, 3), (3,))


def dualquat_to_transform(dq: torch.Tensor) -> torch.Tensor:
    assert dq.shape[-1] == 8
    return _run(_OP_DQ2T, dq, (8,), None, None, (4, 4))


def dualquat_multiply(dq1: torch.Tensor, dq2: torch.Tensor) -> torch.Tensor:
    assert dq1.shape[-1] == 8
    assert dq2.shape[-1] == 8
    return _run(_OP_DQMUL, dq1, (8,), dq2, (8,), (8,))


def dualquat_invert(dq: torch.Tensor) -> torch.Tensor:
    assert dq.shape[-1] == 8
    return _run(_OP_DQINV, dq, (8,), None, None, (8,))


def point_to_dualquat(p: torch.Tensor) -> torch.Tensor:
    assert p.shape[-1] == 3
    return _run(_OP_P2DQ, p, (3,), None, None, (8,))
