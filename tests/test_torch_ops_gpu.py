"""torch.ops.aurdf.* (csrc/torch_ops.cpp, the thin torch extension over the C ABI) against the ctypes
wrappers of the same entry points and against torch autograd of the restated reference expressions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import os
    from autourdf_b200 import torch_ops
    if not os.path.exists(torch_ops._PATH):
        import __graft_entry__ as g
        g.build()
    return torch_ops.load()


def test_icp_sweep_op_equals_ctypes_path(ops):
    from autourdf_b200 import cluster_icp as ci, synth
    b = synth.make_config("wx200", n_frames=4)
    d = ci.batch_to_device(b)
    ms = int(np.diff(b.src_off).max())
    r = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"], max_src_per_tile=ms)
    T, world, corr, fit, rmse, iters, ntgt, status = ops.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"],
                                                                   d["box_off"], d["init_T"], max_src_per_tile=ms)
    assert int(status[0]) == 0
    assert torch.equal(T, r.T) and torch.equal(world, r.world) and torch.equal(corr, r.corr)
    assert torch.equal(iters, r.iters) and torch.equal(ntgt, r.ntgt) and torch.equal(fit, r.fitness) and torch.equal(rmse, r.rmse)
    # no mask (plain registration_icp), float32 storage
    T2, *_ = ops.icp_sweep(d["src"].float(), d["src_off"], d["tgt"].float(), d["tgt_off"], d["tile_frame"], None, None, d["init_T"],
                           max_src_per_tile=ms, max_iter=3)
    assert torch.isfinite(T2).all()


def test_se3_apply_and_chamfer_ops_autograd(ops):
    from autourdf_b200.mlp_reg import calculate_pc
    from autourdf_b200.chamfer import chamfer_distance
    g = torch.Generator().manual_seed(0)
    sizes = [40, 1, 77, 130]
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device="cuda")
    for dt in (torch.float32, torch.float64):
        xyz = torch.randn(sum(sizes), 3, generator=g, dtype=dt).cuda().requires_grad_(True)
        T = torch.randn(len(sizes), 4, 4, generator=g, dtype=dt).cuda().requires_grad_(True)
        out = ops.se3_apply(xyz, off, T)
        parts = torch.split(xyz, sizes)
        ref = torch.cat([p @ T[k][:3, :3].T + T[k][:3, 3] for k, p in enumerate(parts)])      # mlp_reg.py:155-170
        tol = 1e-5 if dt == torch.float32 else 1e-12
        assert (out - ref).abs().max().item() <= tol
        w = torch.randn_like(ref)
        ga = torch.autograd.grad((out * w).sum(), (xyz, T))
        gb = torch.autograd.grad((ref * w).sum(), (xyz, T))
        assert (ga[0] - gb[0]).abs().max().item() <= tol * 10 and (ga[1] - gb[1]).abs().max().item() <= tol * 200
        assert torch.equal(out.detach(), torch.cat(calculate_pc([p.detach() for p in parts], T.detach())))
    x = (torch.randn(2, 600, 3, generator=g) * 0.2).cuda().requires_grad_(True)
    y = (torch.randn(2, 750, 3, generator=g) * 0.2).cuda()
    for norm in (1, 2):
        la = ops.chamfer_distance(x, y, norm)
        lb, _ = chamfer_distance(x, y, norm=norm)
        assert la.item() == lb.item()
        ga, = torch.autograd.grad(la, x)
        gb, = torch.autograd.grad(lb, x)
        assert (ga - gb).abs().max().item() <= 1e-7


def test_dq_ops_match_wrappers_and_reference_expressions(ops):
    from autourdf_b200 import dq_func as D
    from oracle import pt3d_torch as P
    from scipy.spatial.transform import Rotation
    n = 32
    R = torch.tensor(Rotation.random(n, random_state=3).as_matrix(), dtype=torch.float64).cuda()
    T = torch.eye(4, dtype=torch.float64).repeat(n, 1, 1).cuda()
    T[:, :3, :3] = R
    T[:, :3, 3] = torch.randn(n, 3, dtype=torch.float64).cuda()
    dq = torch.randn(n, 8, dtype=torch.float64).cuda()
    q = torch.randn(n, 4, dtype=torch.float64).cuda()
    pairs = [(ops.transform_to_dualquat, D.transform_to_dualquat, P.transform_to_dualquat, T),
             (ops.dualquat_to_transform, D.dualquat_to_transform, P.dualquat_to_transform, dq),
             (ops.quaternion_to_matrix, D.quaternion_to_matrix, P.quaternion_to_matrix, q),
             (ops.matrix_to_quaternion, D.matrix_to_quaternion, P.matrix_to_quaternion, R)]
    for op, wrap, ref, x in pairs:
        a, b_, c = x.clone().requires_grad_(True), x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya, yb, yc = op(a), wrap(b_), ref(c)
        assert torch.equal(ya, yb)
        assert (ya - yc).abs().max().item() <= 1e-10
        w = torch.randn_like(yc)
        (ya * w).sum().backward(); (yb * w).sum().backward(); (yc * w).sum().backward()
        assert torch.equal(a.grad, b_.grad)
        assert (a.grad - c.grad).abs().max().item() <= 1e-9 * max(1.0, c.grad.abs().max().item())
    # by operator code: dualquat_multiply (binary), dualquat_to_rot_trans (two outputs)
    out, empty = ops.dq_op(8, dq, torch.roll(dq, 1, 0).contiguous())
    assert empty.numel() == 0 and torch.equal(out, D.dualquat_multiply(dq, torch.roll(dq, 1, 0)))
    R2, t2 = ops.dq_op(6, dq)
    Rw, tw = D.dualquat_to_rot_trans(dq)
    assert torch.equal(R2, Rw) and torch.equal(t2, tw)
    with pytest.raises(RuntimeError):
        ops.se3_apply(torch.zeros(3, 3), torch.tensor([0, 3], dtype=torch.int32), torch.eye(4)[None])   # CPU tensors: no fallback


def test_nn_l2_op(ops, oracle):
    rng = np.random.default_rng(2)
    qn, tn = rng.normal(size=(300, 3)), rng.normal(size=(500, 3))
    t = lambda a, dt=torch.float64: torch.tensor(a, dtype=dt, device="cuda")
    idx, d2 = ops.nn_l2(t(qn), t([0, 300], torch.int32), t(tn), t([0, 500], torch.int32))
    oi, od = oracle.nn_batch(qn, tn, True)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(d2.cpu().numpy(), od)
