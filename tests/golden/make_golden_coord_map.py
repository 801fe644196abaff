"""Generate tests/golden/coord_map.npz by executing the REFERENCE's own pose bookkeeping:
``CoordMap.load_matrix`` / ``CoordMap.coord_dist_map`` (PointCloud/coord_map.py:186-307) and
``save_pc_npz`` / ``load_pc_npz`` (PointCloud/helper_functions.py:10-21).

Build container only (needs /root/reference):   python tests/golden/make_golden_coord_map.py

coord_map.py imports open3d, matplotlib, pytorch3d, roma and the reference's URDF / mesh modules
at module level; none of the third-party ones is installable here.  Stand-ins are installed in
``sys.modules`` first: empty modules for what the two methods never touch, the pytorch3d
``matrix_to_quaternion`` restated in torch (same as make_golden.py), and ``roma`` restated in
torch from its published algorithms (oracle/coord_map_oracle.py lists them).  The vectors
therefore pin what the reference itself wrote -- file naming and slicing, the xyz+quaternion
layout, which scratch matrix is overwritten when, the three lambdas, the row-wise norms, the
stacking order -- on top of the restated roma pieces, which stay "parity unpinned".
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AURDF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def _install_roma():
    def rotmat_to_unitquat(R):
        R = R.reshape(-1, 3, 3)
        n = R.shape[0]
        dec = torch.empty((n, 4), dtype=R.dtype)
        dec[:, :3] = R.diagonal(dim1=1, dim2=2)
        dec[:, -1] = dec[:, :3].sum(axis=1)
        choices = dec.argmax(axis=1)
        quat = torch.empty((n, 4), dtype=R.dtype)
        ind = torch.nonzero(choices != 3, as_tuple=True)[0]
        i = choices[ind]
        j = (i + 1) % 3
        k = (j + 1) % 3
        quat[ind, i] = 1 - dec[ind, -1] + 2 * R[ind, i, i]
        quat[ind, j] = R[ind, j, i] + R[ind, i, j]
        quat[ind, k] = R[ind, k, i] + R[ind, i, k]
        quat[ind, 3] = R[ind, k, j] - R[ind, j, k]
        ind = torch.nonzero(choices == 3, as_tuple=True)[0]
        quat[ind, 0] = R[ind, 2, 1] - R[ind, 1, 2]
        quat[ind, 1] = R[ind, 0, 2] - R[ind, 2, 0]
        quat[ind, 2] = R[ind, 1, 0] - R[ind, 0, 1]
        quat[ind, 3] = 1 + dec[ind, -1]
        return quat / torch.norm(quat, dim=1)[:, None]

    def unitquat_to_rotvec(quat):
        quat = quat.reshape(-1, 4).clone()
        quat[quat[:, 3] < 0] *= -1
        half = torch.atan2(torch.norm(quat[:, :3], dim=1), quat[:, 3])
        angle = 2 * half
        small = torch.abs(angle) <= 1e-3
        scale = torch.empty(len(quat), dtype=quat.dtype)
        scale[small] = 2 + angle[small] ** 2 / 12 + 7 * angle[small] ** 4 / 2880
        scale[~small] = angle[~small] / torch.sin(half[~small])
        return scale[:, None] * quat[:, :3]

    def rotmat_to_rotvec(R):
        batch = R.shape[:-2]
        return unitquat_to_rotvec(rotmat_to_unitquat(R)).reshape(batch + (3,))

    def rotvec_to_unitquat(v):
        v = v.reshape(-1, 3)
        angle = torch.norm(v, dim=1)
        small = angle <= 1e-3
        scale = torch.empty(len(v), dtype=v.dtype)
        scale[small] = 0.5 - angle[small] ** 2 / 48 + angle[small] ** 4 / 3840
        scale[~small] = torch.sin(angle[~small] / 2) / angle[~small]
        return torch.cat([scale[:, None] * v, torch.cos(angle / 2)[:, None]], 1)

    def rotvec_geodesic_distance(v1, v2):
        q1, q2 = rotvec_to_unitquat(v1), rotvec_to_unitquat(v2)
        d = 4.0 * torch.asin(0.5 * torch.min(torch.norm(q2 - q1, dim=-1), torch.norm(q2 + q1, dim=-1)))
        return d.reshape(v1.shape[:-1])

    def rotmat_geodesic_distance(R1, R2, clamping=1.0):
        return 2.0 * torch.asin(torch.clamp_max(torch.norm(R2 - R1, dim=[-1, -2]) / (2.0 * np.sqrt(2.0)), clamping))

    roma = types.ModuleType("roma")
    roma.rotmat_to_rotvec = rotmat_to_rotvec
    roma.rotmat_geodesic_distance = rotmat_geodesic_distance
    roma.utils = types.SimpleNamespace(rotvec_geodesic_distance=rotvec_geodesic_distance)
    sys.modules["roma"] = roma


def _install_empty(*names):
    for n in names:
        m = types.ModuleType(n)
        m.__getattr__ = lambda name: (lambda *a, **k: None)   # any imported symbol resolves to a no-op
        sys.modules[n] = m


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}; golden vectors can only be generated in the build container")
    import make_golden as G   # the pytorch3d stand-in lives there
    G._install_pytorch3d()
    _install_roma()
    _install_empty("open3d", "matplotlib", "matplotlib.pyplot", "compute_joints", "visualize", "link")
    sys.path.insert(0, os.path.join(REF, "PointCloud"))
    import coord_map as ref_cm            # reference module
    import helper_functions as ref_hf     # reference module

    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(2024)
    T, K = 7, 9
    # a pose track like the registration output: smooth per-cluster motion, two clusters rigidly tied,
    # one static, one with a half-turn between two frames, one with a tiny (series-branch) rotation
    mats = np.tile(np.eye(4), (T, K, 1, 1))
    base_R = Rotation.random(K, random_state=3).as_matrix()
    base_t = rng.uniform(-0.3, 0.3, size=(K, 3))
    axes = rng.normal(size=(K, 3))
    axes /= np.linalg.norm(axes, axis=1)[:, None]
    rate = np.deg2rad(rng.uniform(2, 9, size=K))
    vel = rng.normal(scale=0.01, size=(K, 3))
    rate[2], vel[2] = 0.0, 0.0            # static cluster
    axes[4], rate[4], vel[4] = axes[3], rate[3], vel[3]   # moves like cluster 3
    rate[5] = 1e-5                        # |rotvec| below roma's 1e-3 series threshold
    for t in range(T):
        for k in range(K):
            ang = rate[k] * t + (np.pi if (k == 6 and t >= 4) else 0.0)
            mats[t, k, :3, :3] = Rotation.from_rotvec(axes[k] * ang).as_matrix() @ base_R[k]
            mats[t, k, :3, 3] = base_t[k] + vel[k] * t
    bbox = 0.83

    out = {"matrices": mats, "bounding_box": bbox}
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "matrix"))
        os.makedirs(os.path.join(d, "cluster"))
        for t in range(T):
            np.save(os.path.join(d, "matrix", f"{t:04}.npy"), mats[t])
        cm = object.__new__(ref_cm.CoordMap)          # __init__ needs open3d + .ply files; set the fields it would
        cm.data_path = d + "/"
        rot, m2 = ref_cm.CoordMap.load_matrix(cm, 0, T)
        out["load_matrix_rot"], out["load_matrix_matrices"] = rot, m2
        rot_s, _ = ref_cm.CoordMap.load_matrix(cm, 2, 5)
        out["load_matrix_rot_2_5"] = rot_s
        cm.coords, cm.matrices, cm.num_coords, cm.bounding_box = rot, m2, K, bbox
        for diff in (True, False):
            cmap, smap = ref_cm.CoordMap.coord_dist_map(cm, diff=diff)
            out[f"map_diff{int(diff)}"], out[f"sum_diff{int(diff)}"] = cmap, smap
        # file format round trip through the reference's helpers
        segs = [rng.normal(size=(n, 3)) for n in (5, 1, 12, 0, 7, 3, 4, 2, 9, 6, 8)]   # 11 clusters: key order '0','1','10','2',...
        path = os.path.join(d, "cluster", "0000.npz")
        ref_hf.save_pc_npz(segs, path)
        z = np.load(path)
        out["npz_keys"] = np.array(list(z.keys()))
        back = ref_hf.load_pc_npz(path)
        out["npz_sizes_in"] = np.array([s.shape[0] for s in segs])
        out["npz_sizes_back"] = np.array([s.shape[0] for s in back])
        out["npz_concat_in"] = np.concatenate(segs)
        out["npz_concat_back"] = np.concatenate(back)
    np.savez_compressed(os.path.join(HERE, "coord_map.npz"), **out)
    print("golden written: coord_map.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
