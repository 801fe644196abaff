"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/aurdf.h declares,
and validates arguments without touching a device; host-side packing logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not (os.path.exists(os.path.join(ROOT, "autourdf_b200", "libaurdf.so")) and
            os.path.exists(os.path.join(ROOT, "autourdf_b200", "libaurdf_torch.so"))):
        g.build()
    from autourdf_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "aurdf.h")).read()
    declared = set(re.findall(r"AURDF_API[^;(]*?\b(aurdf_\w+)\s*\(", hdr))
    assert len(declared) >= 14
    L = C.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in aurdf.h but not exported"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert lib.lib().aurdf_version() == 100


def test_argument_validation_without_a_device(lib):
    L = lib.lib()
    assert L.aurdf_icp_workspace_bytes(10, 1000, 5000) > 10 * 6 * 8 + 3 * 5000 * 8
    assert L.aurdf_icp_workspace_bytes(-1, 0, 0) == 0
    # max_corr_dist <= 0 is rejected like open3d does, before any CUDA call
    one = np.zeros(64, dtype=np.float64)
    args = [one.ctypes.data, 1, one.ctypes.data, one.ctypes.data, one.ctypes.data, one.ctypes.data, one.ctypes.data, 0,
            one.ctypes.data, one.ctypes.data, 1, 1, 1, 1.2, 0.0, 10, 1e-6, 1e-6, 0] + [one.ctypes.data] * 7 + \
           [one.ctypes.data, 1 << 20, 16, None, None]
    assert L.aurdf_icp_sweep(*args) == lib.EINVAL
    assert b"max_corr_dist" in L.aurdf_last_error_string()
    assert L.aurdf_dq_op(99, None, None, None, None, 1, 0, None) == lib.EINVAL
    assert L.aurdf_dq_op(0, None, None, None, None, 0, 0, None) == lib.OK          # n == 0: nothing to do
    assert L.aurdf_se3_apply(None, None, None, 0, 0, 0, None, None) == lib.OK
    assert L.aurdf_nn_l2(None, None, None, None, 1, 0, 0, None, None, None) == lib.OK
    assert L.aurdf_icp_sweep_launches() in (4, 5)   # 5 with the small-tile kernel (default)


def test_missing_library_fails_loudly(lib, monkeypatch, tmp_path):
    monkeypatch.setattr(lib, "_LIB", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(lib.AurdfError, match="no CPU fallback"):
        lib.lib()


def test_synth_batch_layout():
    from autourdf_b200 import synth
    b = synth.make_config("wx200", n_frames=3)
    assert b.n_frames == 2 and b.n_tiles == 2 * b.n_clusters
    assert b.src_off[-1] == b.src.shape[0] and b.box_off[-1] == b.box.shape[0] and b.tgt_off[-1] == b.tgt.shape[0]
    assert b.box.dtype == np.float32 and b.src.dtype == np.float64
    assert np.array_equal(b.tgt, b.tgt.astype(np.float32).astype(np.float64))     # float32-representable clouds
    assert np.array_equal(b.init_T, b.init_T.astype(np.float32).astype(np.float64))
    assert np.unique(b.tgt[b.tgt_off[0]:b.tgt_off[1]], axis=0).shape[0] == b.tgt_off[1]   # no duplicate points
    s = b.frame_slice(1, 2)
    assert s.n_tiles == b.n_clusters and s.tgt.shape[0] == b.tgt_off[2] - b.tgt_off[1]
    assert np.array_equal(s.src, b.src[b.src_off[b.n_clusters]:])
    b2 = synth.make_config("wx200", n_frames=3)
    assert np.array_equal(b.tgt, b2.tgt) and np.array_equal(b.src, b2.src)          # deterministic


def test_torch_extension_loads_and_registers_every_operator(lib):
    """the thin torch extension (csrc/torch_ops.cpp) loads without a GPU, registers the operators SURVEY 8(b) lists,
    and refuses CPU tensors (no fallback behind it)"""
    import torch
    from autourdf_b200 import torch_ops
    ops = torch_ops.load()
    for name in ("icp_sweep", "se3_apply", "nn_l2", "dq_op", "transform_to_dualquat", "dualquat_to_transform",
                 "quaternion_to_matrix", "matrix_to_quaternion", "chamfer_distance"):
        assert hasattr(ops, name), name
    assert "Tensor? box" in str(ops.icp_sweep.default._schema)
    with pytest.raises(RuntimeError):
        ops.dualquat_to_transform(torch.zeros(2, 8))
    with pytest.raises(RuntimeError):
        ops.chamfer_distance(torch.zeros(1, 4, 3), torch.zeros(1, 5, 3), 1)
