"""Seeded synthetic articulated-robot sequences shaped like the reference's scans.

The reference ships no data (``data/`` is git-ignored) and its simulator needs
pybullet + OpenGL, so every fixture and bench input comes from here.  The statistics
mirror /root/reference/Sim/sim_data.py: per-point noise sigma 5e-4 (:343), per-frame
translation noise sigma 0.01 for frames > 0 (:337), joint step 4..8 degrees
(:417, step_size 4 at :544), 10 frames per sequence (:545), metres.

A robot is a tree of box links with revolute joints.  Each frame is an independent
i.i.d. surface sample (no point identity across frames, like re-rendered scans), cast
to float32 and widened back to float64 -- the common input of oracle and GPU.
Frame 0 of sequence 0 is segmented with k-means++ (cluster_icp.py:67) and every
cluster gets the local frame (I, centroid) (cluster_icp.py:86-99); all sequences
reuse those clusters as the ICP source (mlp_reg.py:250-253 and the quirk noted in
SURVEY.md section 3).  The init pose of tile (frame f -> f+1, cluster k) is the
ground-truth cluster pose at frame f (float32-valued, like ``step_m`` at
mlp_reg.py:322); the box source is the float32 prediction ``init @ local``
(``pred_pcd_np``, mlp_reg.py:121).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# name -> (N points / frame, K clusters, sequences, frames / sequence, dof, config id)
CONFIGS = {
    "wx200": dict(n_points=1024, n_clusters=10, n_seq=1, n_frames=10, dof=5, cid=1),
    "wx200_5": dict(n_points=2048, n_clusters=20, n_seq=5, n_frames=10, dof=5, cid=2),
    "franka": dict(n_points=4096, n_clusters=30, n_seq=5, n_frames=10, dof=7, cid=3),
    "allegro_hand": dict(n_points=8192, n_clusters=48, n_seq=5, n_frames=10, dof=16, cid=4),
}


def _rot(axis, ang):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _se3(R=None, t=None):
    T = np.eye(4)
    if R is not None:
        T[:3, :3] = R
    if t is not None:
        T[:3, 3] = t
    return T


@dataclass
class Robot:
    """links[i] = (parent, joint origin in parent frame (3,), joint axis (3,), box half-extents (3,),
    box centre in link frame (3,)).  Link 0 is the fixed base (parent -1, no joint)."""
    links: list

    @property
    def dof(self):
        return len(self.links) - 1

    def fk(self, q):
        """world pose (4,4) of every link for joint vector q (dof,)"""
        G = [None] * len(self.links)
        for i, (parent, origin, axis, _, _) in enumerate(self.links):
            if parent < 0:
                G[i] = _se3(t=origin)
            else:
                G[i] = G[parent] @ _se3(_rot(axis, q[i - 1]), origin)
        return G


def make_robot(dof: int, rng) -> Robot:
    axes = [np.array([0, 0, 1.0]), np.array([0, 1.0, 0]), np.array([1.0, 0, 0])]
    if dof < 12:  # serial arm: base + dof links along +z with alternating axes
        L = 0.55 / (dof + 1)
        links = [(-1, np.array([0, 0, -0.3]), None, np.array([0.05, 0.05, L / 2]), np.array([0, 0, L / 2]))]
        for j in range(dof):
            w = 0.045 - 0.02 * j / max(dof - 1, 1)
            links.append((j, np.array([0, 0, L]), axes[(j + 1) % 3 if j else 0],
                          np.array([w * (0.8 + 0.4 * rng.random()), w * (0.8 + 0.4 * rng.random()), L / 2]),
                          np.array([0, 0, L / 2])))
        return Robot(links)
    # hand: palm + n_f fingers x 4 joints
    n_f = dof // 4
    links = [(-1, np.array([0, 0, -0.1]), None, np.array([0.05 * n_f / 2, 0.015, 0.05]), np.array([0, 0, 0.05]))]
    for f in range(n_f):
        x = (f - (n_f - 1) / 2) * 0.045
        parent = 0
        for j in range(4):
            origin = np.array([x, 0, 0.1]) if j == 0 else np.array([0, 0, 0.045])
            axis = axes[1] if j == 0 else axes[2]
            links.append((parent, origin, axis, np.array([0.011, 0.011, 0.0225]), np.array([0, 0, 0.0225])))
            parent = len(links) - 1
    return Robot(links)


def sample_surface(robot: Robot, G, n, rng):
    """n i.i.d. points on the box surfaces (area-weighted), world frame; returns (pts, link id)"""
    face_l, face_ax, face_s, areas = [], [], [], []
    for li, (_, _, _, h, c) in enumerate(robot.links):
        for ax in range(3):
            u, v = (ax + 1) % 3, (ax + 2) % 3
            for s in (-1.0, 1.0):
                areas.append(4 * h[u] * h[v])
                face_l.append(li); face_ax.append(ax); face_s.append(s)
    areas = np.asarray(areas)
    face_l, face_ax, face_s = np.asarray(face_l), np.asarray(face_ax), np.asarray(face_s)
    H = np.stack([l[3] for l in robot.links])      # half extents per link
    Cc = np.stack([l[4] for l in robot.links])     # box centre per link
    Rw = np.stack([g[:3, :3] for g in G])
    tw = np.stack([g[:3, 3] for g in G])
    pick = rng.choice(areas.size, size=n, p=areas / areas.sum())
    uv = rng.uniform(-1, 1, size=(n, 2))
    lid = face_l[pick]
    ax = face_ax[pick]
    h = H[lid]
    rows = np.arange(n)
    p = np.empty((n, 3))
    p[rows, ax] = face_s[pick] * h[rows, ax]
    p[rows, (ax + 1) % 3] = uv[:, 0] * h[rows, (ax + 1) % 3]
    p[rows, (ax + 2) % 3] = uv[:, 1] * h[rows, (ax + 2) % 3]
    p += Cc[lid]
    pts = np.einsum("nij,nj->ni", Rw[lid], p) + tw[lid]
    return pts, lid


def _cloud(robot, q, n, rng, frame_noise):
    G = robot.fk(q)
    pts, lid = sample_surface(robot, G, n, rng)
    pts = pts + rng.normal(0.0, 5e-4, size=pts.shape) + frame_noise
    pts = pts.astype(np.float32)
    # reject duplicate points (exact NN ties are unpinnable): re-jitter until unique
    for _ in range(8):
        _, first = np.unique(pts, axis=0, return_index=True)
        if first.size == pts.shape[0]:
            break
        dup = np.setdiff1d(np.arange(pts.shape[0]), first)
        pts[dup] = (pts[dup].astype(np.float64) + rng.normal(0.0, 5e-4, size=(dup.size, 3))).astype(np.float32)
    return pts.astype(np.float64), lid, G


@dataclass
class SweepBatch:
    """Packed (frame, cluster) tiles of one or more sequences -- the layout of the C ABI."""
    src: np.ndarray          # (sum n_s, 3) f64   local clusters, tile-major
    src_off: np.ndarray      # (B+1,) i32
    tgt: np.ndarray          # (sum M, 3) f64     one target cloud per frame transition
    tgt_off: np.ndarray      # (F+1,) i32
    tile_frame: np.ndarray   # (B,) i32
    box: np.ndarray          # (sum n_b, 3) f32   predicted world clusters (AABB source)
    box_off: np.ndarray      # (B+1,) i32
    init_T: np.ndarray       # (B, 4, 4) f64, float32-valued
    n_frames: int = 0        # F (frame transitions)
    n_clusters: int = 0      # K
    meta: dict = field(default_factory=dict)

    @property
    def n_tiles(self):
        return int(self.tile_frame.shape[0])

    def frame_slice(self, f0, f1):
        """sub-batch holding frame transitions [f0, f1) (tiles are frame-major)"""
        tiles = np.nonzero((self.tile_frame >= f0) & (self.tile_frame < f1))[0]
        b0, b1 = (int(tiles[0]), int(tiles[-1]) + 1) if tiles.size else (0, 0)
        s0, s1 = int(self.src_off[b0]), int(self.src_off[b1])
        x0, x1 = int(self.box_off[b0]), int(self.box_off[b1])
        t0, t1 = int(self.tgt_off[f0]), int(self.tgt_off[f1])
        return SweepBatch(self.src[s0:s1], self.src_off[b0:b1 + 1] - s0, self.tgt[t0:t1],
                          self.tgt_off[f0:f1 + 1] - t0, self.tile_frame[b0:b1] - f0, self.box[x0:x1],
                          self.box_off[b0:b1 + 1] - x0, self.init_T[b0:b1], f1 - f0, self.n_clusters,
                          dict(self.meta))


    def frame_select(self, frames):
        """sub-batch holding the given frame transitions (ascending list) and their tiles, plus the indices of
        those tiles in this batch: the unit of interleaved (round-robin) frame sharding"""
        frames = np.asarray(frames, dtype=np.int64)
        assert frames.size == 0 or np.all(np.diff(frames) > 0)
        keep = np.isin(self.tile_frame, frames)
        tiles = np.nonzero(keep)[0]
        z = np.zeros(1, dtype=np.int32)
        if tiles.size == 0:
            return SweepBatch(self.src[:0], z, self.tgt[:0], z.copy(), self.tile_frame[:0], self.box[:0], z.copy(),
                              self.init_T[:0], 0, self.n_clusters, dict(self.meta)), tiles
        remap = np.full(int(self.tgt_off.shape[0]) - 1, -1, dtype=np.int32)
        remap[frames] = np.arange(frames.size, dtype=np.int32)
        cat = lambda arr, off, idx: np.concatenate([arr[off[i]:off[i + 1]] for i in idx])
        offs = lambda off, idx: np.concatenate([[0], np.cumsum([off[i + 1] - off[i] for i in idx])]).astype(np.int32)
        return SweepBatch(cat(self.src, self.src_off, tiles), offs(self.src_off, tiles), cat(self.tgt, self.tgt_off, frames),
                          offs(self.tgt_off, frames), remap[self.tile_frame[tiles]], cat(self.box, self.box_off, tiles),
                          offs(self.box_off, tiles), self.init_T[tiles], int(frames.size), self.n_clusters, dict(self.meta)), tiles

    def tile_slice(self, b0, b1):
        """sub-batch holding tiles [b0, b1) and (a copy of) every frame they refer to: the unit of
        cluster-level sharding, where each GPU holds a replica of the frame's target cloud"""
        b0, b1 = int(b0), int(b1)
        if b1 <= b0:
            z = np.zeros(1, dtype=np.int32)
            return SweepBatch(self.src[:0], z, self.tgt[:0], z, self.tile_frame[:0], self.box[:0], z.copy(),
                              self.init_T[:0], 0, self.n_clusters, dict(self.meta))
        frames = np.unique(self.tile_frame[b0:b1])                      # sorted
        remap = np.full(int(self.tgt_off.shape[0]) - 1, -1, dtype=np.int32)
        remap[frames] = np.arange(frames.size, dtype=np.int32)
        tgt = np.concatenate([self.tgt[self.tgt_off[f]:self.tgt_off[f + 1]] for f in frames])
        tgt_off = np.zeros(frames.size + 1, dtype=np.int32)
        tgt_off[1:] = np.cumsum([self.tgt_off[f + 1] - self.tgt_off[f] for f in frames])
        s0, s1 = int(self.src_off[b0]), int(self.src_off[b1])
        x0, x1 = int(self.box_off[b0]), int(self.box_off[b1])
        return SweepBatch(self.src[s0:s1], self.src_off[b0:b1 + 1] - s0, tgt, tgt_off, remap[self.tile_frame[b0:b1]],
                          self.box[x0:x1], self.box_off[b0:b1 + 1] - x0, self.init_T[b0:b1], int(frames.size),
                          self.n_clusters, dict(self.meta))


def kmeans_frame0(pts, k, seed):
    """cluster_icp.py:67 -- sklearn k_means(init='k-means++'); returns labels"""
    from sklearn.cluster import k_means
    _, labels, _ = k_means(pts, n_clusters=k, init="k-means++", n_init=1, random_state=seed)
    return labels


def make_batch(n_points=2048, n_clusters=20, n_seq=5, n_frames=10, dof=5, cid=2, seed=None, tile_repeat=1):
    """Build the packed batch of all frame transitions of ``n_seq`` sequences.

    ``tile_repeat`` > 1 replicates the whole set of sequences with fresh noise seeds (used by
    the bench to make a workload larger than L2)."""
    src_l, box_l, tgt_l, init_l, tile_frame = [], [], [], [], []
    clusters_local = None
    F = 0
    base_seed = cid * 1000 if seed is None else seed
    rng0 = np.random.default_rng(base_seed)
    robot = make_robot(dof, rng0)
    q_home = rng0.uniform(-0.4, 0.4, size=robot.dof)
    for rep in range(tile_repeat):
        for s in range(n_seq):
            rng = np.random.default_rng(base_seed + s + 100 * rep)
            # piece-wise linear joint trajectory, |dq| in U(4,8) degrees per step
            q = q_home.copy()
            sign = rng.choice([-1.0, 1.0], size=robot.dof)
            frames, poses = [], []
            for f in range(n_frames):
                noise = rng.normal(0.0, 0.01, size=3) if f > 0 else np.zeros(3)
                pts, lid, G = _cloud(robot, q, n_points, rng, noise)
                frames.append((pts, lid, [_se3(t=noise) @ g for g in G]))
                flip = rng.random(robot.dof) < 0.15
                sign = np.where(flip, -sign, sign)
                q = q + sign * np.deg2rad(4.0 * (1.0 + rng.random(robot.dof)))
            if clusters_local is None:
                pts0, lid0, G0 = frames[0]
                labels = kmeans_frame0(pts0, n_clusters, base_seed)
                clusters_local, cl_link, cl_T0 = [], [], []
                for k in range(n_clusters):
                    m = labels == k
                    ck = pts0[m]
                    cen = ck.mean(axis=0)
                    clusters_local.append(ck - cen)          # inverse of (I, centroid): cluster_icp.py:96-98
                    cl_link.append(int(np.bincount(lid0[m]).argmax()))
                    cl_T0.append(_se3(t=cen))
                G_home_inv = [np.linalg.inv(g) for g in G0]
            for f in range(n_frames - 1):
                _, _, Gf = frames[f]
                tgt_l.append(frames[f + 1][0])
                for k in range(n_clusters):
                    l = cl_link[k]
                    init = (Gf[l] @ G_home_inv[l] @ cl_T0[k]).astype(np.float32)
                    init[3] = (0, 0, 0, 1)
                    pred = clusters_local[k].astype(np.float32) @ init[:3, :3].T + init[:3, 3]
                    src_l.append(clusters_local[k])
                    box_l.append(pred.astype(np.float32))
                    init_l.append(init.astype(np.float64))
                    tile_frame.append(F)
                F += 1

    def _pack(arrs, dt):
        off = np.zeros(len(arrs) + 1, dtype=np.int32)
        off[1:] = np.cumsum([a.shape[0] for a in arrs])
        return (np.concatenate(arrs).astype(dt) if arrs else np.zeros((0, 3), dt)), off

    src, src_off = _pack(src_l, np.float64)
    box, box_off = _pack(box_l, np.float32)
    tgt, tgt_off = _pack(tgt_l, np.float64)
    return SweepBatch(src, src_off, tgt, tgt_off, np.asarray(tile_frame, dtype=np.int32), box, box_off,
                      np.asarray(init_l, dtype=np.float64).reshape(-1, 4, 4), F, n_clusters,
                      dict(n_points=n_points, n_clusters=n_clusters, n_seq=n_seq * tile_repeat,
                           n_frames=n_frames, dof=dof, seed=base_seed))


def make_config(name: str, **over) -> SweepBatch:
    cfg = dict(CONFIGS[name])
    cfg.update(over)
    return make_batch(**cfg)
