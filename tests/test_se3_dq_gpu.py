"""GPU parity of the SE(3) apply kernels and the dual-quaternion library against the numpy
oracle and the golden vectors produced by the reference's own dq_func.py / mlp_reg.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

F32_TOL = 2e-6   # float32 ops, same expression order: a few ulp of O(1) values
F64_TOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "dq_func.npz"))


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_dq_func_golden(gold, dt):
    from autourdf_b200 import dq_func as D
    tol = F32_TOL if dt == "f32" else F64_TOL
    g = lambda k: gold[f"{dt}_{k}"]
    R, t, q, dq_raw, p = (_dev(g(k)) for k in ("in_R", "in_t", "in_q", "in_dq_raw", "in_p"))
    T = D.transform_from_rot_trans(R, t)
    dq = D.transform_to_dualquat(T)
    dq_b = D.transform_to_dualquat(torch.roll(T, 1, 0))
    qq, tt = D.dualquat_to_quat_trans(dq)
    R2, t2 = D.dualquat_to_rot_trans(dq_raw)
    got = dict(transform_from_rot_trans=T, quaternion_conjugate=D.quaternion_conjugate(q),
               quat_trans_to_dualquat=D.quat_trans_to_dualquat(q, t), rot_trans_to_dualquat=D.rot_trans_to_dualquat(R, t),
               transform_to_dualquat=dq, dualquat_to_quat_trans_q=qq, dualquat_to_quat_trans_t=tt,
               dualquat_to_rot_trans_R=R2, dualquat_to_rot_trans_t=t2,
               dualquat_to_transform=D.dualquat_to_transform(dq_raw), dualquat_multiply=D.dualquat_multiply(dq, dq_b),
               dualquat_invert=D.dualquat_invert(dq_raw), point_to_dualquat=D.point_to_dualquat(p),
               matrix_to_quaternion=D.matrix_to_quaternion(R), quaternion_to_matrix=D.quaternion_to_matrix(dq_raw[:, :4]))
    for k, v in got.items():
        ref = g(k)
        v = v.cpu().numpy()
        assert v.shape == ref.shape and v.dtype == ref.dtype, k
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(v - ref).max() <= tol * scale * (40 if k in ("dualquat_invert", "dualquat_to_transform", "dualquat_to_rot_trans_R", "quaternion_to_matrix") else 1), (k, np.abs(v - ref).max())


def test_dq_func_vs_numpy_oracle_and_properties():
    from autourdf_b200 import dq_func as D
    from oracle import dq_oracle as O
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(0)
    n = 4096
    Rm = Rotation.random(n, random_state=1).as_matrix()
    t = rng.normal(size=(n, 3))
    T = O.transform_from_rot_trans(Rm, t)
    Td = _dev(T)
    dq = D.transform_to_dualquat(Td)
    assert np.abs(dq.cpu().numpy() - O.transform_to_dualquat(T)).max() <= F64_TOL
    assert np.abs(D.dualquat_to_transform(dq).cpu().numpy() - T).max() <= 1e-12           # round trip
    T2 = O.transform_from_rot_trans(Rotation.random(n, random_state=2).as_matrix(), rng.normal(size=(n, 3)))
    prod = D.dualquat_to_transform(D.dualquat_multiply(dq, D.transform_to_dualquat(_dev(T2))))
    assert np.abs(prod.cpu().numpy() - T @ T2).max() <= 1e-11                              # dq(A) dq(B) <-> A B
    ident = D.dualquat_multiply(dq, D.dualquat_invert(dq)).cpu().numpy()
    assert np.abs(ident - np.array([1, 0, 0, 0, 0, 0, 0, 0.0])).max() <= 1e-12             # dq dq^-1 = 1
    q = D.matrix_to_quaternion(_dev(Rm)).cpu().numpy()
    qs = Rotation.from_matrix(Rm).as_quat()                                                 # scalar-last
    qs = np.concatenate([qs[:, 3:], qs[:, :3]], -1)
    qs = np.where(qs[:, :1] < 0, -qs, qs)
    assert np.abs(q - qs).max() <= 1e-12
    # batch dims broadcast like the reference: (2, n/2, 8) x (n/2, 8)
    a = dq.reshape(2, n // 2, 8)
    b = dq[: n // 2]
    got = D.dualquat_multiply(a, b).cpu().numpy()
    assert np.abs(got - O.dualquat_multiply(a.cpu().numpy(), b.cpu().numpy()[None])).max() <= F64_TOL
    with pytest.raises(AssertionError):
        D.dualquat_invert(dq[:, :7])


def test_calculate_pc_golden_and_autograd(golden_dir):
    from autourdf_b200.mlp_reg import calculate_pc
    z = np.load(os.path.join(golden_dir, "calculate_pc.npz"))
    off = z["off"]
    K = off.shape[0] - 1
    cl = [torch.from_numpy(z["local"][off[k]:off[k + 1]]).cuda() for k in range(K)]
    mats = torch.from_numpy(z["matrices"]).cuda()
    pcs = calculate_pc(cl, mats)
    got = torch.cat(pcs).cpu().numpy()
    assert got.dtype == np.float32 and [p.shape for p in pcs] == [c.shape for c in cl]
    assert np.abs(got - z["world"]).max() <= F32_TOL
    # autograd against the reference expression, float64
    cl64 = [c.double().requires_grad_(True) for c in cl]
    m64 = mats.double().requires_grad_(True)
    w = torch.randn(sum(c.shape[0] for c in cl), 3, dtype=torch.float64, device="cuda")
    (torch.cat(calculate_pc(cl64, m64)) * w).sum().backward()
    g_cl = [c.grad.clone() for c in cl64]
    g_m = m64.grad.clone()
    cl_r = [c.detach().clone().requires_grad_(True) for c in cl64]
    m_r = m64.detach().clone().requires_grad_(True)
    ref = torch.cat([c @ m_r[i][:3, :3].T + m_r[i][:3, 3] for i, c in enumerate(cl_r)])   # mlp_reg.py:168
    (ref * w).sum().backward()
    for a, b in zip(g_cl, cl_r):
        assert (a - b.grad).abs().max().item() <= 1e-12
    assert (g_m - m_r.grad).abs().max().item() <= 1e-10


def test_to_local_matches_reference_expression():
    from autourdf_b200.mlp_reg import to_local
    from oracle import dq_oracle as O
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(5)
    K = 7
    pts = [rng.normal(size=(rng.integers(0, 300), 3)) for _ in range(K)]
    mats = O.transform_from_rot_trans(Rotation.random(K, random_state=3).as_matrix(), rng.normal(size=(K, 3)))
    got = to_local(pts, mats)
    for k in range(K):
        assert got[k].shape == pts[k].shape
        if pts[k].shape[0]:
            assert np.abs(got[k] - O.to_local(pts[k], mats[k])).max() <= 1e-12


def test_nn_l2_vs_oracle(oracle):
    from autourdf_b200 import _lib
    rng = np.random.default_rng(9)
    groups = [(300, 500), (0, 10), (17, 0), (2500, 3100), (1, 1)]
    q = [rng.normal(size=(a, 3)) for a, _ in groups]
    t = [rng.normal(size=(b, 3)) for _, b in groups]
    qoff = np.zeros(len(groups) + 1, np.int32); qoff[1:] = np.cumsum([a for a, _ in groups])
    toff = np.zeros(len(groups) + 1, np.int32); toff[1:] = np.cumsum([b for _, b in groups])
    for dt in (np.float64, np.float32):
        qa = np.concatenate(q).astype(dt); ta = np.concatenate(t).astype(dt)
        idx = torch.empty(int(qoff[-1]), dtype=torch.int32, device="cuda")
        d2 = torch.empty(int(qoff[-1]), dtype=torch.float64, device="cuda")
        L = _lib.lib()
        dq, dqo, dt_, dto = _dev(qa), _dev(qoff), _dev(ta), _dev(toff)    # keep the device buffers alive
        _lib.check(L.aurdf_nn_l2(_lib.ptr(dq), _lib.ptr(dqo), _lib.ptr(dt_), _lib.ptr(dto),
                                 _lib.F32 if dt == np.float32 else _lib.F64, len(groups), int(qoff[-1]),
                                 _lib.ptr(idx), _lib.ptr(d2), _lib.current_stream()))
        torch.cuda.synchronize()
        idx = idx.cpu().numpy(); d2 = d2.cpu().numpy()
        for g, (a, b) in enumerate(groups):
            sl = slice(qoff[g], qoff[g + 1])
            if a == 0:
                continue
            if b == 0:
                assert (idx[sl] == -1).all()
                continue
            oi, od = oracle.nn_batch(qa[sl].astype(np.float64), ta[toff[g]:toff[g + 1]].astype(np.float64), True)
            assert np.array_equal(idx[sl], oi) and np.array_equal(d2[sl], od)


def test_resample_cluster_vs_oracle_and_sklearn():
    """seeded Lloyd k-means + local frames; labels identical to the oracle (which is
    pinned against sklearn.cluster.k_means in the CPU tests), local clusters to 1e-12"""
    import warnings
    from sklearn.cluster import k_means
    from autourdf_b200 import synth
    from autourdf_b200.mlp_reg import resample_cluster
    from oracle import kmeans_oracle as KO
    b = synth.make_config("wx200_5", n_seq=1)
    K = b.n_clusters
    for f in (0, 3, 8):
        cloud = b.tgt[b.tgt_off[f]:b.tgt_off[f + 1]]
        mats = b.init_T[f * K:(f + 1) * K]
        d = {}
        got = resample_cluster(cloud, 0, K, mats, _details=d)
        ref, labels = KO.resample_cluster(cloud, K, mats)
        assert np.array_equal(d["labels"], labels)
        assert [g.shape for g in got] == [r.shape for r in ref]
        for g, r in zip(got, ref):
            if r.shape[0]:
                assert np.abs(g - r).max() <= 1e-12
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            c, l, inertia = k_means(cloud, n_clusters=K, init=mats[:, :3, 3], n_init=1)
        assert np.array_equal(d["labels"], l) and np.abs(d["centers"] - c).max() <= 1e-12
        assert abs(d["inertia"] - inertia) <= 1e-9 * inertia
    # a seed far from the cloud -> empty cluster -> re-seeded with the farthest point, like sklearn
    mats = b.init_T[:K].copy()
    mats[3, :3, 3] = [5.0, 5.0, 5.0]
    cloud = b.tgt[b.tgt_off[0]:b.tgt_off[1]]
    d = {}
    got = resample_cluster(cloud, 0, K, mats, _details=d)
    _, labels = KO.resample_cluster(cloud, K, mats)
    assert np.array_equal(d["labels"], labels)

    class _Seg:                       # the reference passes its Segments object and a frame index
        class _PC:
            def __init__(self, p): self.points = p
        def __init__(self, clouds): self.pc_list = [self._PC(c) for c in clouds]
    got2 = resample_cluster(_Seg([cloud, cloud]), 1, K, mats)
    assert all(np.array_equal(a, b_) for a, b_ in zip(got, got2))
    with pytest.raises(NotImplementedError):
        resample_cluster(cloud, 0, K, mats, normal=True)


def _rand_dq_inputs(n, dt, seed=0):
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(seed)
    R = torch.tensor(Rotation.random(n, random_state=seed).as_matrix(), dtype=dt)
    t = torch.randn(n, 3, generator=g, dtype=dt)
    T = torch.eye(4, dtype=dt).repeat(n, 1, 1)
    T[:, :3, :3], T[:, :3, 3] = R, t
    q = torch.randn(n, 4, generator=g, dtype=dt)
    dq = torch.randn(n, 8, generator=g, dtype=dt)
    return dict(R=R, t=t, T=T, q=q, q2=torch.randn(n, 4, generator=g, dtype=dt), dq=dq,
                dq2=torch.randn(n, 8, generator=g, dtype=dt), p=torch.randn(n, 3, generator=g, dtype=dt))


def test_dq_func_backward_matches_finite_differences():
    """every operator of aurdf_dq_op_bwd against torch.autograd.gradcheck (float64, central differences)"""
    from autourdf_b200 import dq_func as D
    x = {k: v.cuda() for k, v in _rand_dq_inputs(6, torch.float64, seed=3).items()}
    r = lambda a: a.clone().requires_grad_(True)
    cases = [(D.transform_from_rot_trans, (x["R"], x["t"])), (D.quaternion_conjugate, (x["q"],)),
             (D.quat_trans_to_dualquat, (x["q"], x["t"])), (D.rot_trans_to_dualquat, (x["R"], x["t"])),
             (D.transform_to_dualquat, (x["T"],)), (D.dualquat_to_quat_trans, (x["dq"],)),
             (D.dualquat_to_rot_trans, (x["dq"],)), (D.dualquat_to_transform, (x["dq"],)),
             (D.dualquat_multiply, (x["dq"], x["dq2"])), (D.dualquat_invert, (x["dq"],)), (D.point_to_dualquat, (x["p"],)),
             (D.quaternion_raw_multiply, (x["q"], x["q2"])), (D.quaternion_invert, (x["q"],)),
             (D.quaternion_to_matrix, (x["q"],)), (D.matrix_to_quaternion, (x["R"],))]
    for fn, args in cases:
        assert torch.autograd.gradcheck(fn, tuple(r(a) for a in args), eps=1e-6, atol=1e-6, rtol=1e-5), fn.__name__


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_dq_func_backward_matches_torch_autograd_of_the_reference_expressions(dt):
    """the two pairs train() differentiates through (mlp_reg.py:60-66 `--r q`, :78-84 `--r dq`): gradients of
    the CUDA operators vs torch autograd through the restated reference expressions, incl. broadcasting"""
    from autourdf_b200 import dq_func as D
    from oracle import pt3d_torch as P
    x = {k: v.cuda() for k, v in _rand_dq_inputs(48, dt, seed=7).items()}
    x["R"][1] = torch.diag(torch.tensor([1.0, -1.0, -1.0], dtype=dt)).cuda()      # other arg-max branches
    x["R"][2] = torch.diag(torch.tensor([-1.0, 1.0, -1.0], dtype=dt)).cuda()
    x["T"][1, :3, :3], x["T"][2, :3, :3] = x["R"][1], x["R"][2]
    tol = 2e-4 if dt == torch.float32 else 1e-10
    pairs = [(D.matrix_to_quaternion, P.matrix_to_quaternion, "R"), (D.quaternion_to_matrix, P.quaternion_to_matrix, "q"),
             (D.transform_to_dualquat, P.transform_to_dualquat, "T"), (D.dualquat_to_transform, P.dualquat_to_transform, "dq")]
    for ours, ref, key in pairs:
        a, b = x[key].clone().requires_grad_(True), x[key].clone().requires_grad_(True)
        ya, yb = ours(a), ref(b)
        w = torch.randn_like(yb)
        assert (ya - yb).abs().max().item() <= tol, ours.__name__
        (ya * w).sum().backward()
        (yb * w).sum().backward()
        scale = max(1.0, b.grad.abs().max().item())
        assert (a.grad - b.grad).abs().max().item() <= tol * scale, f"{ours.__name__}: gradient differs"
    # the composition of the `--r dq` branch with a parameter in between, and a broadcast operand
    w8 = torch.randn(8, dtype=dt, device="cuda", requires_grad=True)
    T = D.dualquat_to_transform(D.transform_to_dualquat(x["T"]) * w8)
    T.square().sum().backward()
    w8r = w8.detach().clone().requires_grad_(True)
    P.dualquat_to_transform(P.transform_to_dualquat(x["T"]) * w8r).square().sum().backward()
    assert (w8.grad - w8r.grad).abs().max().item() <= tol * max(1.0, w8r.grad.abs().max().item()) * 10
    qa = x["q"][:1].clone().requires_grad_(True)                                  # (1,4) broadcast against (48,4)
    D.quaternion_raw_multiply(qa, x["q2"]).sum().backward()
    qb = x["q"][:1].clone().requires_grad_(True)
    P.quaternion_raw_multiply(qb, x["q2"]).sum().backward()
    assert (qa.grad - qb.grad).abs().max().item() <= tol * 50
