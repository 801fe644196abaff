"""CPU property tests (hypothesis + seeded numpy): the oracle's algebra, and the soundness of the
float32 nearest-neighbour certificate that icp_small_kernel / icp_tiles_kernel rely on.

The certificate test is an emulation, not the kernel: it repeats the kernel's float32 arithmetic
(coordinates about the tile's first target, rounded to float32; packed distance; key = distance
bits with the low 10 bits replaced by the target index; best / second-best key; tau) in numpy and
checks the claim the kernels' bit-exactness rests on -- *whenever the certificate accepts the
float32 winner, it is the float64 argmin with the reference's operation order*.  The kernels take
the exact float64 scan for everything the certificate rejects, so a sound certificate is all that
is needed (autourdf_b200/csrc/icp_small.cu, "float32 pre-filter").

Two more emulations guard arithmetic that only runs on the GPU: the lane split / read-ahead index
arithmetic of icp_small_kernel (every target pair scanned exactly once per point, no shared-memory
read past the initialised pair array) and its Newton-on-SO(3) pose fit (quaternion accumulation,
unnormalised update, chord finish, certificate) against umeyama through the C oracle."""
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


# ------------------------------------------------------------------ lane split / read-ahead arithmetic of icp_small_kernel
def _lane_split(ns, npairs, NT=128, PAIRS=384):
    """the integer arithmetic at the top of icp_small_kernel (icp_small.cu: S0, rounds, S_tail, nfill)"""
    S0 = 1
    while S0 < 32 and ns * (S0 * 2) <= NT and npairs + 8 * S0 <= PAIRS:
        S0 *= 2
    ppr = NT // S0
    rounds = (ns + ppr - 1) // ppr
    S_tail = S0
    if rounds > 1:
        rem = ns - (rounds - 1) * ppr
        S_tail = 1
        while S_tail < 32 and rem * (S_tail * 2) <= NT and npairs + 8 * S_tail <= PAIRS:
            S_tail *= 2
    return S0, ppr, rounds, S_tail, min(PAIRS, npairs + 4 * max(S0, S_tail))


def test_small_kernel_lane_split_covers_every_pair_and_stays_in_bounds():
    """every (point, target pair) is scanned exactly once, every lane maps to at most one point per
    round, and the software-pipelined scan never reads past the initialised part of the pair array"""
    NT, PAIRS = 128, 384
    for ns in list(range(1, 140)) + [150, 191, 192, 193, 224, 255, 256, 257, 300, 319, 320]:
        for nt in (1, 2, 3, 7, 64, 131, 255, 256, 383, 384, 385, 700, 759, 760):
            npairs = (nt + 1) // 2
            S0, ppr, rounds, S_tail, nfill = _lane_split(ns, npairs, NT, PAIRS)
            assert nfill <= PAIRS and rounds * ppr >= ns
            seen_points = set()
            for r in range(rounds):
                S = S_tail if r == rounds - 1 else S0
                trips2 = ((npairs + S - 1) // S + 1) >> 1
                pts = {}
                for tid in range(NT):
                    i, sub = tid // S + r * ppr, tid & (S - 1)
                    if i >= ns:
                        continue
                    # pairs this lane touches: sub + S * (2 t) and sub + S * (2 t + 1), t < trips2; it also LOADS one trip ahead
                    touched = [sub + S * k for k in range(2 * trips2)]
                    assert sub + S * (2 * trips2 + 1) < nfill, (ns, nt, r, S)
                    pts.setdefault(i, []).extend(touched)
                for i, tl in pts.items():
                    real = sorted(j for j in tl if j < npairs)
                    assert real == list(range(npairs)), (ns, nt, r, i)        # each real pair exactly once
                    assert i not in seen_points
                    seen_points.add(i)
            assert seen_points == set(range(ns))


# ------------------------------------------------------------------ the Newton-on-SO(3) pose fit (emulation of kabsch_rotation_newton4)
def _newton4(sigma):
    A = sigma.copy()
    tol = 1e-16 * (abs(A[0, 0]) + abs(A[1, 1]) + abs(A[2, 2]))
    q = np.array([1.0, 0.0, 0.0, 0.0])
    vv, conv, steps = 1.0, False, 0
    adj = hr = None
    for _ in range(8):
        k = np.array([A[2, 1] - A[1, 2], A[0, 2] - A[2, 0], A[1, 0] - A[0, 1]])
        if np.abs(k).max() <= tol:
            conv = True
            break
        if not vv < 1e-10:
            S = 0.5 * (A + A.T)
            G = np.trace(S) * np.eye(3) - S
            det = np.linalg.det(G)
            adj, hr = np.linalg.inv(G) * det, 0.5 / det
        v = hr * (adj @ k)
        vv = float(v @ v)
        if not vv < 1.0:
            return None
        w, x, y, z = q
        q = np.array([w - (x * v[0] + y * v[1] + z * v[2]), x + (w * v[0] + (y * v[2] - z * v[1])),
                      y + (w * v[1] + (z * v[0] - x * v[2])), z + (w * v[2] + (x * v[1] - y * v[0]))])
        steps += 1
        if vv < 1e-16:
            conv = True
            break
        A = (1.0 - vv) * A + 2.0 * (np.outer(v, v @ A) - np.cross(v, A.T).T)     # (1 + v.v) E^T A, column by column
        if vv < 1e-13:
            conv = True
            break
    if not conv:
        return None
    tr = np.trace(A)
    e1 = 1e-9 * tr
    if not (A[0, 0] > e1 and A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0] > e1 * tr and np.linalg.det(A) > e1 * tr * tr):
        return None
    w, x, y, z = q
    s = 2.0 / (q @ q)
    R = np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                  [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                  [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])
    return R, steps


def test_newton_pose_fit_equals_umeyama_or_declines(oracle):
    """the small-tile kernel's rotation fit: whenever it certifies a result, that result is umeyama's
    U V^T to rounding; reflections and rank-deficient covariances are declined (the kernel then runs
    the Jacobi SVD)"""
    rng = np.random.default_rng(11)
    n_ok = 0
    for mag in (1.0, 0.3, 1e-1, 1e-2, 1e-3, 1e-5, 1e-8):
        for _ in range(60):
            n = int(rng.integers(4, 80))
            P = rng.normal(size=(n, 3)) * rng.uniform(0.005, 0.1, size=3)
            Rt = Rotation.from_rotvec(rng.normal(size=3) / np.sqrt(3) * mag).as_matrix()
            Q = P @ Rt.T + rng.normal(scale=5e-4, size=P.shape)
            sigma = (Q - Q.mean(0)).T @ (P - P.mean(0)) / n
            ref = oracle.kabsch(P, Q, np.arange(n, dtype=np.int32))[:3, :3]
            r = _newton4(sigma)
            if r is None:
                assert mag >= 0.3          # only large rotations may be handed to the fallback
                continue
            R, steps = r
            n_ok += 1
            assert np.abs(R - ref).max() <= 1e-12 and np.abs(R @ R.T - np.eye(3)).max() <= 1e-14
            if mag <= 1e-3:
                assert steps <= 3          # late ICP iterations: one full step + a short finish (or two)
    assert n_ok > 300
    mirror = np.diag([1.0, 1.0, -1.0]) * 1e-4                      # det < 0: umeyama flips an axis; Newton must decline
    assert _newton4(mirror) is None
    rank1 = np.outer([1.0, 2.0, 3.0], [0.5, -1.0, 2.0]) * 1e-5     # collinear matches
    assert _newton4(rank1) is None


# ------------------------------------------------------------------ icp_small2_kernel: nearest-neighbour cache
def _cache_state(P, Q):
    """what a scan of icp_small2_kernel leaves per point: winner j1, anchor (float32 about Q[0]), and
    rho = lower bound on the distance from the anchor position to every other target (0 = no bound)"""
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
        cert = (m2lo - m1hi > np.float32(2.0) * tau) & (m2hi < np.inf)
        rho = np.where((m2lo > np.float32(128.0) * dl * dl) & (m2lo - tau > 0),
                       np.sqrt(np.maximum(m2lo - tau, np.float32(0))) * np.float32(0.9999), np.float32(0)).astype(np.float32)
    if nt == 1:
        cert = np.ones_like(cert)
        rho = np.full_like(rho, 1e30)
    rho = np.where(cert, rho, np.float32(0))          # an uncertified point keeps no bound
    return j1, f32, rho, aq


def _cache_hit(Pnew, Q, j1, anchor32, rho, aq):
    """the kernel's cache test for points that moved to Pnew (float32 arithmetic as written there)"""
    o = Q[0]
    f = (Pnew - o).astype(np.float32)
    d = Pnew - Q[j1]
    D1 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    e = f - anchor32
    ex, ey, ez = (e[:, k].astype(np.float64) for k in range(3))
    s = (ex * ex).astype(np.float32)
    s = (ey * ey + s.astype(np.float64)).astype(np.float32)
    s = (ez * ez + s.astype(np.float64)).astype(np.float32)
    dele = np.sqrt(s)
    amag = np.maximum(aq, np.abs(f).max(axis=1)).astype(np.float32)
    lhs = (np.sqrt(D1.astype(np.float32)) + dele) * np.float32(1.0001) + np.float32(1e-6) * (amag + dele)
    return (rho > 0) & (lhs < rho)


@pytest.mark.parametrize("kind", ["random", "bisector", "duplicates"])
def test_nn_cache_certificate_is_sound(kind):
    """whenever the cache test of icp_small2_kernel keeps the previous winner after a move, that winner
    is the float64 argmin at the new position (reference operation order, lowest index on ties)"""
    rng = np.random.default_rng({"random": 11, "bisector": 12, "duplicates": 13}[kind])
    checked = hits = 0
    for trial in range(120):
        nt = int(rng.choice([1, 2, 3, 17, 64, 131, 300, 760]))
        P, Q = _tile(rng, kind, nt, 192)
        j1, anchor32, rho, aq = _cache_state(P, Q)
        ext = np.abs(Q - Q[0]).max() + 1e-6
        for scale in (1e-7, 1e-5, 1e-3, 3e-2, 0.3):      # moves from far below to far above the target spacing
            step = rng.normal(size=P.shape) * (scale * ext)
            if kind != "random":                          # move straight at a rival target now and then
                rival = Q[rng.integers(0, nt, P.shape[0])]
                step = np.where(rng.random((P.shape[0], 1)) < 0.5, (rival - P) * rng.uniform(0, 1.2, (P.shape[0], 1)), step)
            Pn = P + step
            hit = _cache_hit(Pn, Q, j1, anchor32, rho, aq)
            ex = _exact_argmin(Pn, Q)
            wrong = hit & (j1 != ex)
            assert not wrong.any(), f"{kind}: cache kept a wrong neighbour (trial {trial}, n_t {nt}, move {scale})"
            checked += hit.size
            hits += int(hit.sum())
    assert hits / checked > (0.25 if kind == "random" else 0.05), f"cache useless: {hits / checked:.3f}"


def test_small2_home_slot_mapping_is_a_bijection():
    """icp_small2_kernel deals the source points round-robin to its four warps: slot <-> point"""
    s2p = lambda t: (t & ~127) + 4 * (t & 31) + ((t >> 5) & 3)
    p2s = lambda i: (i & ~127) + 32 * (i & 3) + ((i & 127) >> 2)
    for n in (128, 256):
        assert sorted(s2p(t) for t in range(n)) == list(range(n))
        assert all(p2s(s2p(t)) == t for t in range(n))
    for ns in (1, 5, 20, 113, 128, 129, 200, 256):     # every warp's home points fill its first lanes
        for r in range((ns + 127) // 128):
            for w in range(4):
                act = [128 * r + 4 * l + w < ns for l in range(32)]
                k = sum(act)
                assert act == [True] * k + [False] * (32 - k)
                left = ns - 128 * r - w
                assert k == (0 if left <= 0 else min(32, (left + 3) >> 2))
                assert (0 if left <= 0 else min(8, (left + 15) >> 4)) == (k + 3) // 4


# ------------------------------------------------------------------ grid-pruned search of icp_grid_kernel (icp_grid.cu)
def _grid_build(Q):
    """build_grid of icp_grid.cu in numpy float32: origin = first target, about one target per cell by volume, at most
    64 cells per axis and n_t cells; returns (f32 targets about the origin, cell of every target, g0, h, inv_h, G, amax)"""
    f = (Q - Q[0]).astype(np.float32)
    lo, hi = f.min(0), f.max(0)
    e = (hi - lo).astype(np.float32)
    emax, nt = np.float32(e.max()), Q.shape[0]
    G, h = np.array([1, 1, 1]), np.float32(1.0)
    if emax > 0:
        fl = np.float32(1e-3) * emax
        h = np.float32(np.cbrt(np.maximum(e[0], fl) * np.maximum(e[1], fl) * np.maximum(e[2], fl) / np.float32(nt)))
        h = np.maximum(h, emax / np.float32(64.0))
        while True:
            G = np.clip((e / h).astype(np.int64) + 1, 1, 64)
            if G.prod() <= nt:
                break
            h = np.float32(h * np.float32(1.1))
    inv_h = np.float32(1.0) / h
    cell = lambda v: np.clip(np.floor((v - lo) * inv_h).astype(np.int64), 0, G - 1)
    return f, cell(f), lo, h, inv_h, G, np.float32(np.abs(f).max()), cell


def _grid_query(p32, f, cells, lo, h, G, aq, cell, R=None):
    """one query of pass_walk: the block (3x3x3 of the query's cell, widened to the cube of half-side R and then until
    it holds the ball through the best target found), float32 best / second best, face distance, certificate.
    Returns (winner index or -1, certified)"""
    amag = np.float32(max(aq, np.abs(p32).max()))
    dl = np.float32(4.0) * U32 * amag
    eps = np.float32(2e-5) * h + np.float32(8.0) * U32 * amag
    c = cell(p32)
    bl, bh = np.maximum(c - 1, 0), np.minimum(c + 1, G - 1)

    def widen(bl, bh, R):
        nl, nh = cell(p32 - R), cell(p32 + R)
        grew = bool((nl < bl).any() or (nh > bh).any())
        return np.minimum(bl, nl), np.maximum(bh, nh), grew

    if R is not None:
        bl, bh, _ = widen(bl, bh, np.float32(R) + np.float32(64.0) * dl + eps)
    ring = 1
    while True:
        inside = ((cells >= bl) & (cells <= bh)).all(1)
        idx = np.nonzero(inside)[0]
        d = ((p32 - f[idx]) ** 2)
        d = (d[:, 2] + (d[:, 1] + d[:, 0])).astype(np.float32) if idx.size else np.zeros(0, np.float32)   # fmaf chain ~ float32 sum
        whole = bool((bl == 0).all() and (bh == G - 1).all())
        if whole:
            break
        if idx.size:
            m1 = d.min()
            bl, bh, grew = widen(bl, bh, np.sqrt(m1) * np.float32(1.01) + np.float32(64.0) * dl + eps)
            if not grew:
                break
        else:
            ring *= 2
            bl, bh = np.maximum(bl - ring, 0), np.minimum(bh + ring, G - 1)
    if idx.size == 0:
        return -1, False
    order = np.argsort(d, kind="stable")
    m1, k1 = d[order[0]], idx[order[0]]
    m2 = d[order[1]] if idx.size > 1 else np.float32(np.inf)
    # face distance: nearest face of the block that is not a face of the grid, deflated by eps
    fb = np.float32(np.inf)
    for a in range(3):
        if bl[a] > 0:
            fb = min(fb, np.float32(p32[a] - (lo[a] + np.float32(bl[a]) * h)))
        if bh[a] < G[a] - 1:
            fb = min(fb, np.float32((lo[a] + np.float32(bh[a] + 1) * h) - p32[a]))
    if np.isfinite(fb):
        fb = max(np.float32(fb - eps), np.float32(0.0))
    m2e = min(m2, np.float32(fb * fb))
    with np.errstate(invalid="ignore", over="ignore"):
        tau = np.float32(16.0) * (dl * np.sqrt(m2e) * np.float32(1.001) + dl * dl + U32 * m2e)
        ok = (f.shape[0] == 1 and not np.isfinite(m2e)) or (np.isfinite(m2e) and m2e - m1 > np.float32(2.0) * tau)
    return int(k1), bool(ok)


@pytest.mark.parametrize("kind", ["random", "bisector", "duplicates", "outside"])
def test_grid_search_certificate_is_sound(kind):
    """numpy emulation of the grid search of icp_grid_kernel on adversarial tiles: whenever the float32 block scan +
    face-distance certificate accepts a winner, it is the float64 brute-force argmin over ALL targets (the kernel
    sends everything else to the exact float64 path); queries far outside the grid, on bisector planes, on
    duplicated targets, with and without a radius from a previous winner"""
    rng = np.random.default_rng({"random": 11, "bisector": 12, "duplicates": 13, "outside": 14}[kind])
    checked = accepted = 0
    for trial in range(40):
        nt = int(rng.choice([2, 17, 64, 300, 900, 2500]))
        P, Q = _tile(rng, "random" if kind == "outside" else kind, nt, 48)
        if kind == "outside":
            P = P + rng.normal(scale=3.0 * np.ptp(Q, axis=0).max(), size=3)        # the whole cluster away from the targets
        f, cells, lo, h, inv_h, G, aq, cell = _grid_build(Q)
        ex = _exact_argmin(P, Q)
        for i in range(P.shape[0]):
            p32 = (P[i] - Q[0]).astype(np.float32)
            # half the queries carry the radius through a previous winner (a random target: any target is an upper bound)
            R = None
            if i % 2:
                jprev = int(rng.integers(0, nt))
                R = np.float32(np.sqrt(((P[i] - Q[jprev]) ** 2).sum()) * 1.01)
            k1, ok = _grid_query(p32, f, cells, lo, h, G, aq, cell, R)
            checked += 1
            if ok:
                accepted += 1
                assert k1 == ex[i], f"{kind}: certified a wrong neighbour (trial {trial}, n_t {nt}, point {i}): {k1} vs {ex[i]}"
    frac = accepted / checked
    assert frac > (0.97 if kind in ("random", "outside") else 0.4), f"{kind}: only {frac:.3f} of the queries certified"
