"""Generate tests/golden/*.npz by executing the REFERENCE's own Python for the hot path.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference modules import open3d and pytorch3d, which are not installable here.  This
script installs two small stand-ins into ``sys.modules`` before importing them:

  * ``open3d``      -- PointCloud / Vector3dVector / registration_icp backed by the CPU
                       oracle (oracle/icp_oracle.c), i.e. the restated open3d semantics;
  * ``pytorch3d``   -- the four ``transforms`` functions dq_func.py imports, restated in
                       torch from the published pytorch3d 0.7.7 algorithm.

So the vectors pin everything the reference itself wrote -- ``masked_icp``'s box / mask /
ori / output handling (PointCloud/cluster_icp.py:118-191), all 11 ``dq_func`` functions
(PointCloud/dq_func.py:4-257) and ``calculate_pc`` (PointCloud/mlp_reg.py:155-170) -- on
top of the restated third-party pieces, which stay "parity unpinned" (DESIGN.md).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AURDF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import icp_oracle as O  # noqa: E402


# ------------------------------------------------------------------ open3d stand-in
def _install_open3d():
    o3d = types.ModuleType("open3d")

    class Vector3dVector:
        def __init__(self, a):
            self.a = np.array(a, dtype=np.float64).reshape(-1, 3)

        def __array__(self, dtype=None, copy=None):
            return self.a

    class PointCloud:
        def __init__(self, pts=None):
            self.points = pts if pts is not None else Vector3dVector(np.zeros((0, 3)))

        def paint_uniform_color(self, c):
            return self

        def transform(self, T):
            self.points = Vector3dVector(O.transform_pts(np.asarray(T, dtype=np.float64), self.points.a))
            return self

    class _Result:
        pass

    def registration_icp(source, target, max_correspondence_distance, init=np.eye(4), estimation_method=None,
                         criteria=None):
        r = O.icp_p2p(source.points.a, target.points.a, max_correspondence_distance,
                      np.asarray(init, dtype=np.float64), max_iter=criteria.max_iteration if criteria else 30,
                      rel_fit=criteria.relative_fitness if criteria else 1e-6,
                      rel_rmse=criteria.relative_rmse if criteria else 1e-6)
        out = _Result()
        out.transformation = r["T"]
        out.fitness = r["fitness"]
        out.inlier_rmse = r["rmse"]
        i = np.nonzero(r["corr"] >= 0)[0]
        out.correspondence_set = np.stack([i, r["corr"][i]], 1)
        return out

    class ICPConvergenceCriteria:
        def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
            self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration

    class TransformationEstimationPointToPoint:
        def __init__(self, with_scaling=False):
            assert not with_scaling

    o3d.geometry = types.SimpleNamespace(PointCloud=PointCloud)
    o3d.utility = types.SimpleNamespace(Vector3dVector=Vector3dVector)
    o3d.pipelines = types.SimpleNamespace(registration=types.SimpleNamespace(
        registration_icp=registration_icp, ICPConvergenceCriteria=ICPConvergenceCriteria,
        TransformationEstimationPointToPoint=TransformationEstimationPointToPoint))
    o3d.visualization = types.SimpleNamespace(draw_geometries=lambda *a, **k: None)
    o3d.io = types.SimpleNamespace()
    sys.modules["open3d"] = o3d


# ------------------------------------------------------------------ pytorch3d stand-in
def _install_pytorch3d():
    from oracle.pt3d_torch import (matrix_to_quaternion, quaternion_invert, quaternion_raw_multiply,  # noqa: F401
                                   quaternion_to_matrix)

    p3d = types.ModuleType("pytorch3d")
    tr = types.ModuleType("pytorch3d.transforms")
    tr.quaternion_raw_multiply = quaternion_raw_multiply
    tr.quaternion_invert = quaternion_invert
    tr.quaternion_to_matrix = quaternion_to_matrix
    tr.matrix_to_quaternion = matrix_to_quaternion
    for name in ("matrix_to_euler_angles", "euler_angles_to_matrix", "matrix_to_rotation_6d", "rotation_6d_to_matrix"):
        setattr(tr, name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError(name)))
    loss = types.ModuleType("pytorch3d.loss")
    loss.chamfer_distance = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("chamfer_distance"))
    p3d.transforms, p3d.loss = tr, loss
    sys.modules.update({"pytorch3d": p3d, "pytorch3d.transforms": tr, "pytorch3d.loss": loss})


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}; golden vectors can only be generated in the build container")
    _install_open3d()
    _install_pytorch3d()
    sys.path.insert(0, os.path.join(REF, "PointCloud"))
    import cluster_icp as ref_icp   # reference module
    import dq_func as ref_dq        # reference module
    import mlp_reg as ref_reg       # reference module (for calculate_pc)

    from autourdf_b200 import synth

    # ---------------- masked_icp: small seeded sweep through the reference's own function
    b = synth.make_batch(n_points=600, n_clusters=6, n_seq=1, n_frames=3, dof=4, cid=7)
    K = b.n_clusters
    out = {}
    for f in range(b.n_frames):
        tiles = range(f * K, (f + 1) * K)
        cl_local = [b.src[b.src_off[t]:b.src_off[t + 1]] for t in tiles]
        cl_world = [b.box[b.box_off[t]:b.box_off[t + 1]] for t in tiles]          # float32, like pred_pcd_np
        cloud = b.tgt[b.tgt_off[f]:b.tgt_off[f + 1]]
        mats = b.init_T[f * K:(f + 1) * K].astype(np.float32)                      # float32, like step_m_np
        for ori in (False, True):
            np.random.seed(0)
            w, m = ref_icp.masked_icp(cl_local, cl_world, cloud, mats, False, ori=ori)
            out[f"f{f}_ori{int(ori)}_T"] = m
            out[f"f{f}_ori{int(ori)}_world"] = np.concatenate(w)
        # float64 predicted clusters and non-default scale / threshold
        w, m = ref_icp.masked_icp(cl_local, [c.astype(np.float64) for c in cl_world], cloud, mats, False, ori=False,
                                  scale=1.5, th=0.02)
        out[f"f{f}_f64box_T"] = m
        out[f"f{f}_f64box_world"] = np.concatenate(w)
    np.savez_compressed(os.path.join(HERE, "masked_icp.npz"), src=b.src, src_off=b.src_off, tgt=b.tgt, tgt_off=b.tgt_off,
                        tile_frame=b.tile_frame, box=b.box, box_off=b.box_off, init_T=b.init_T, K=K, **out)

    # ---------------- dq_func: all 11 functions, float32 (the live dtype, mlp_reg.py:78-84) and float64
    rng = np.random.default_rng(42)
    from scipy.spatial.transform import Rotation
    n = 64
    dqout = {}
    for dt, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        Rm = Rotation.random(n, random_state=5).as_matrix()
        Rm[0] = np.eye(3)
        Rm[1] = np.diag([1.0, -1.0, -1.0])     # 180 deg about x: exercises the argmax branch
        Rm[2] = np.diag([-1.0, 1.0, -1.0])
        Rm[3] = np.diag([-1.0, -1.0, 1.0])
        t = rng.normal(size=(n, 3))
        R_t, t_t = torch.tensor(Rm, dtype=tdt), torch.tensor(t, dtype=tdt)
        T = ref_dq.transform_from_rot_trans(R_t, t_t)
        q = sys.modules["pytorch3d.transforms"].matrix_to_quaternion(R_t)
        dq = ref_dq.transform_to_dualquat(T)
        dq_b = ref_dq.transform_to_dualquat(torch.roll(T, 1, 0))
        dq_raw = torch.tensor(rng.normal(size=(n, 8)), dtype=tdt)                # un-normalised input
        p = torch.tensor(rng.normal(size=(n, 3)), dtype=tdt)
        qq, tt = ref_dq.dualquat_to_quat_trans(dq)
        R2, t2 = ref_dq.dualquat_to_rot_trans(dq_raw)
        res = dict(in_R=R_t, in_t=t_t, in_q=q, in_dq_raw=dq_raw, in_p=p,
                   transform_from_rot_trans=T, quaternion_conjugate=ref_dq.quaternion_conjugate(q),
                   quat_trans_to_dualquat=ref_dq.quat_trans_to_dualquat(q, t_t),
                   rot_trans_to_dualquat=ref_dq.rot_trans_to_dualquat(R_t, t_t), transform_to_dualquat=dq,
                   dualquat_to_quat_trans_q=qq, dualquat_to_quat_trans_t=tt,
                   dualquat_to_rot_trans_R=R2, dualquat_to_rot_trans_t=t2,
                   dualquat_to_transform=ref_dq.dualquat_to_transform(dq_raw),
                   dualquat_multiply=ref_dq.dualquat_multiply(dq, dq_b),
                   dualquat_invert=ref_dq.dualquat_invert(dq_raw), point_to_dualquat=ref_dq.point_to_dualquat(p),
                   matrix_to_quaternion=q,
                   quaternion_to_matrix=sys.modules["pytorch3d.transforms"].quaternion_to_matrix(dq_raw[:, :4]))
        for k, v in res.items():
            dqout[f"{dt}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "dq_func.npz"), **dqout)

    # ---------------- calculate_pc (mlp_reg.py:155-170), float32 torch
    cl = [torch.tensor(b.src[b.src_off[k]:b.src_off[k + 1]], dtype=torch.float32) for k in range(K)]
    mats = torch.tensor(b.init_T[:K], dtype=torch.float32)
    pcs = ref_reg.calculate_pc(cl, mats)
    np.savez_compressed(os.path.join(HERE, "calculate_pc.npz"), local=np.concatenate([c.numpy() for c in cl]),
                        off=b.src_off[:K + 1], matrices=mats.numpy(), world=np.concatenate([p.numpy() for p in pcs]))
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
