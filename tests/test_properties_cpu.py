"""CPU property tests (hypothesis + seeded numpy): the oracle's algebra, and the soundness of the
float32 nearest-neighbour certificate that icp_small_kernel / icp_tiles_kernel rely on.

The certificate test is an emulation, not the kernel: it repeats the kernel's float32 arithmetic
(coordinates about the tile's first target, rounded to float32; packed distance; key = distance
bits with the low 10 bits replaced by the target index; best / second-best key; tau) in numpy and
checks the claim the kernels' bit-exactness rests on -- *whenever the certificate accepts the
float32 winner, it is the float64 argmin with the reference's operation order*.  The kernels take
the exact float64 scan for everything the certificate rejects, so a sound certificate is all that
is needed (autourdf_b200/csrc/icp_small.cu, "float32 pre-filter")."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from scipy.spatial.transform import Rotation


# ------------------------------------------------------------------ SE(3) / dual quaternions (oracle)
def _rigid(seed, n):
    rng = np.random.default_rng(seed)
    T = np.tile(np.eye(4), (n, 1, 1))
    T[:, :3, :3] = Rotation.random(n, random_state=seed).as_matrix()
    T[:, :3, 3] = rng.normal(size=(n, 3))
    return T


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 10**6))
def test_dualquat_round_trip_product_inverse(seed):
    from oracle import dq_oracle as D
    A, B = _rigid(seed, 8), _rigid(seed + 1, 8)
    dqa, dqb = D.transform_to_dualquat(A), D.transform_to_dualquat(B)
    assert np.abs(D.dualquat_to_transform(dqa) - A).max() <= 1e-12
    assert np.abs(D.dualquat_to_transform(D.dualquat_multiply(dqa, dqb)) - A @ B).max() <= 1e-12
    ident = D.dualquat_multiply(dqa, D.dualquat_invert(dqa))
    assert np.abs(D.dualquat_to_transform(ident) - np.eye(4)).max() <= 1e-12
    q = D.matrix_to_quaternion(A[:, :3, :3])
    assert (q[:, 0] >= 0).all() and np.abs(D.quaternion_to_matrix(q) - A[:, :3, :3]).max() <= 1e-12
    ref = Rotation.from_matrix(A[:, :3, :3]).as_quat()[:, [3, 0, 1, 2]]          # scalar first
    ref *= np.where(ref[:, :1] < 0, -1.0, 1.0)
    assert np.abs(q - ref).max() <= 1e-12


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10**6), st.integers(3, 60))
def test_kabsch_recovers_rigid_motion_and_stays_proper(oracle, seed, n):
    rng = np.random.default_rng(seed)
    P = rng.normal(size=(n, 3))
    T = _rigid(seed, 1)[0]
    Q = P @ T[:3, :3].T + T[:3, 3]
    U = oracle.kabsch(P, Q, np.arange(n, dtype=np.int32))
    if np.linalg.matrix_rank(P - P.mean(0), tol=1e-9) == 3:
        assert np.abs(U - T).max() <= 1e-9
    assert abs(np.linalg.det(U[:3, :3]) - 1.0) <= 1e-9                              # never a reflection
    Qm = Q * np.array([1.0, 1.0, -1.0])                                              # mirrored target
    Um = oracle.kabsch(P, Qm, np.arange(n, dtype=np.int32))
    assert abs(np.linalg.det(Um[:3, :3]) - 1.0) <= 1e-9


@settings(max_examples=10, deadline=None)
@given(st.integers(0, 10**6))
def test_icp_oracle_invariants(oracle, seed):
    """permuting the targets only relabels the correspondences; an unreachable target leaves the pose alone"""
    rng = np.random.default_rng(seed)
    tgt = rng.uniform(-0.1, 0.1, size=(200, 3))
    src = tgt[:80] @ Rotation.from_rotvec(rng.normal(scale=0.02, size=3)).as_matrix() + rng.normal(scale=1e-3, size=3)
    a = oracle.icp_p2p(src, tgt, 1.0, np.eye(4), max_iter=50)
    p = rng.permutation(200)
    b = oracle.icp_p2p(src, tgt[p], 1.0, np.eye(4), max_iter=50)
    assert a["iters"] == b["iters"] and np.array_equal(p[b["corr"]], a["corr"])
    assert np.abs(a["T"] - b["T"]).max() <= 1e-9
    far = oracle.icp_p2p(src, tgt + 10.0, 0.5, np.eye(4), max_iter=50)                # nothing within 0.5
    assert (far["corr"] == -1).all() and np.array_equal(far["T"], np.eye(4)) and far["fitness"] == 0.0


# ------------------------------------------------------------------ the float32 certificate
IDX_MASK = np.uint32(0x3FF)
U32 = np.float32(5.9604645e-8)


def _f32_keys(P, Q):
    """keys of the kernel's float32 scan: (n_s, n_t) uint32, plus aq and the float32 points"""
    o = Q[0]
    q32 = (Q - o).astype(np.float32)
    f32 = (P - o).astype(np.float32)
    d = f32[:, None, :] + (-q32)[None, :, :]                       # float32 add, as FADD2 on negated targets
    dx, dy, dz = (d[..., k].astype(np.float64) for k in range(3))
    acc = (dx * dx).astype(np.float32)                             # FMUL2
    acc = (dy * dy + acc.astype(np.float64)).astype(np.float32)    # FFMA2: product exact in float64, one rounding
    acc = (dz * dz + acc.astype(np.float64)).astype(np.float32)
    keys = (acc.view(np.uint32) & ~IDX_MASK) | np.arange(Q.shape[0], dtype=np.uint32)[None, :]
    return keys, np.abs(q32).max(), f32


def _certified_winner(P, Q):
    keys, aq, f32 = _f32_keys(P, Q)
    nt = Q.shape[0]
    order = np.sort(keys, axis=1)
    m1 = order[:, 0]
    m2 = order[:, 1] if nt > 1 else np.full_like(m1, 0xFFFFFFFF)
    j1 = (m1 & IDX_MASK).astype(np.int64)
    m1hi = (m1 | IDX_MASK).view(np.float32)
    m2lo = (m2 & ~IDX_MASK).view(np.float32)
    m2hi = (m2 | IDX_MASK).view(np.float32)
    amag = np.maximum(aq, np.abs(f32).max(axis=1)).astype(np.float32)
    dl = np.float32(4.0) * U32 * amag
    with np.errstate(invalid="ignore", over="ignore"):
        tau = np.float32(16.0) * (dl * np.sqrt(m2hi) * np.float32(1.001) + dl * dl + U32 * m2hi)
        ok = (m2lo - m1hi > np.float32(2.0) * tau) & (m2hi < np.inf)
    if nt == 1:
        ok = np.ones_like(ok)
    return j1, ok


def _exact_argmin(P, Q):
    """nanoflann / open3d order: ((dx dx) + dy dy) + dz dz in float64, first index on ties"""
    d = P[:, None, :] - Q[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    return d2.argmin(axis=1)


def _tile(rng, kind, nt, ns):
    """targets + source points of one tile, with the near-ties a real run never shows this densely"""
    centre = rng.uniform(-0.5, 0.5, 3)
    ext = 10 ** rng.uniform(-2.5, -0.3)                             # tile extent 3 mm .. 0.5 m
    Q = centre + rng.uniform(-ext, ext, size=(nt, 3))
    if kind == "duplicates" and nt > 3:                             # coincident and nearly coincident targets
        Q[1] = Q[0]
        Q[3] = Q[2] + rng.normal(scale=10 ** rng.uniform(-9, -5), size=3)
    Q = Q.astype(np.float32).astype(np.float64)                     # scans are float32-valued, as in the reference data
    P = centre + rng.uniform(-1.2 * ext, 1.2 * ext, size=(ns, 3))
    if kind in ("bisector", "duplicates") and nt > 1:               # points (almost) equidistant from two targets
        a, b = Q[rng.integers(0, nt, ns)], Q[rng.integers(0, nt, ns)]
        n = b - a
        t = rng.normal(size=(ns, 3))
        t -= (t * n).sum(1, keepdims=True) * n / np.maximum((n * n).sum(1, keepdims=True), 1e-30)
        eps = (10 ** rng.uniform(-12, -2, size=(ns, 1))) * rng.choice([-1.0, 1.0], size=(ns, 1))
        P = 0.5 * (a + b) + 0.3 * ext * t / np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-30) + eps * n
    return P, Q


@pytest.mark.parametrize("kind", ["random", "bisector", "duplicates"])
def test_float32_certificate_is_sound(kind):
    rng = np.random.default_rng({"random": 1, "bisector": 2, "duplicates": 3}[kind])
    checked = accepted = 0
    for trial in range(150):
        nt = int(rng.choice([1, 2, 3, 17, 64, 131, 300, 760]))
        P, Q = _tile(rng, kind, nt, 192)
        j1, ok = _certified_winner(P, Q)
        ex = _exact_argmin(P, Q)
        wrong = ok & (j1 != ex)
        assert not wrong.any(), (f"{kind}: certificate accepted a wrong neighbour in trial {trial} (n_t = {nt}): "
                                 f"point {np.nonzero(wrong)[0][:3]}")
        checked += ok.size
        accepted += int(ok.sum())
    # the filter must also be useful: nearly everything is certified on ordinary data, and even the
    # adversarial sets keep a majority (the rest takes the exact float64 scan in the kernel)
    frac = accepted / checked
    assert frac > (0.995 if kind == "random" else 0.5), f"{kind}: only {frac:.3f} of the points certified"
