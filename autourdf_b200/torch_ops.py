"""``torch.ops.aurdf.*``: the thin torch extension over the C ABI.

``libaurdf_torch.so`` (``csrc/torch_ops.cpp``, built by ``__graft_entry__.build()`` / ``make -C autourdf_b200/csrc
torchext``) registers ``TORCH_LIBRARY(aurdf, ...)`` operators that check device / dtype / contiguity, allocate
outputs and workspace as torch tensors, take ``at::cuda::getCurrentCUDAStream()`` and call ``libaurdf.so``:

    icp_sweep, se3_apply (autograd), nn_l2, dq_op (autograd; 15 operators by code), transform_to_dualquat,
    dualquat_to_transform, quaternion_to_matrix, matrix_to_quaternion (autograd), chamfer_distance (autograd)

The ``ctypes`` wrappers of this package (``cluster_icp``, ``mlp_reg``, ``dq_func``, ``chamfer``) and these operators
are two bindings of the same entry points; there is no CPU fallback behind either.
"""
from __future__ import annotations

import os

import torch

from . import _lib

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libaurdf_torch.so")
_loaded = False


def load():
    """load the extension (once) and return ``torch.ops.aurdf``"""
    global _loaded
    if not _loaded:
        if not os.path.exists(_PATH):
            raise _lib.AurdfError(f"{_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no fallback for the torch operators)")
        _lib.lib()                       # libaurdf.so first (the extension links against it by $ORIGIN)
        torch.ops.load_library(_PATH)
        _loaded = True
    return torch.ops.aurdf
