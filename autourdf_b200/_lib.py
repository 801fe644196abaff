"""ctypes loader of the C-ABI library ``libaurdf.so`` (include/aurdf.h).

There is no CPU fallback: if the library is missing, every operator fails loudly.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C autourdf_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaurdf.so")

F32, F64 = 0, 1
OK, EINVAL, EWORKSPACE, ECUDA, ECAPACITY, ENOMEM = 0, -1, -2, -3, -4, -5

_vp, _i32, _i64, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); every symbol include/aurdf.h declares
SIGNATURES = {
    "aurdf_version": (C.c_int, []),
    "aurdf_last_error_string": (C.c_char_p, []),
    "aurdf_icp_workspace_bytes": (_sz, [_i32, _i64, _i64]),
    "aurdf_icp_sweep": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _i32, _i64, _i32,
                                  _f64, _f64, _i32, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _sz, _i64, _vp, _vp]),
    "aurdf_icp_sweep_launches": (C.c_int, []),
    "aurdf_icp_profile_enable": (C.c_int, [C.c_int]),
    "aurdf_icp_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(_i32)]),
    "aurdf_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "aurdf_ctx_destroy": (None, [_vp]),
    "aurdf_icp_sweep_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _i32, _i32,
                                       _f64, _f64, _i32, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aurdf_ctx_last_copy_bytes": (None, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "aurdf_nn_l2": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i32, _i64, _vp, _vp, _vp]),
    "aurdf_nn_f32_workspace_bytes": (_sz, [_i64]),
    "aurdf_nn_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "aurdf_nn_f32_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, C.c_int, _vp, _vp, _vp]),
    "aurdf_chamfer_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "aurdf_chamfer_workspace_init": (C.c_int, [_vp, _sz, _vp]),
    "aurdf_chamfer_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aurdf_chamfer_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "aurdf_se3_apply": (C.c_int, [_vp, _vp, _vp, _i32, _i64, C.c_int, _vp, _vp]),
    "aurdf_se3_apply_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, C.c_int, _vp, _vp, _vp]),
    "aurdf_se3_to_local": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "aurdf_resample_clusters": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aurdf_dq_op": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _i64, C.c_int, _vp]),
    "aurdf_dq_op_bwd": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_int, _vp]),
    "aurdf_coord_dist_map_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "aurdf_coord_dist_map": (C.c_int, [_vp, _i32, _i32, _f64, _i32, _vp, _vp, _vp, _sz, _vp]),
}

_LIB = None


class AurdfError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise AurdfError(f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                             "`python -c \"import __graft_entry__ as g; g.build()\"` "
                             "(or `make -C autourdf_b200/csrc`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        missing = [n for n in SIGNATURES if not hasattr(L, n)]
        if missing:
            raise AurdfError(f"{LIB_PATH} is stale: missing symbols {missing}; rebuild it")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int, what: str = "aurdf"):
    if rc != OK:
        msg = lib().aurdf_last_error_string().decode(errors="replace")
        raise AurdfError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """device (or host) address of a torch tensor / numpy array, None -> NULL"""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
