"""Second, independent CPU restatement of the hot path in numpy/scipy (float64).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED at the open3d boundary (see icp_oracle.c).
It exists to cross-check ``icp_oracle.c`` with different building blocks: scipy cKDTree
for the nearest neighbour, LAPACK (numpy.linalg.svd) for the 3x3 SVD, numpy boolean
masks exactly as written in /root/reference/PointCloud/cluster_icp.py:133-148.
Poses agree with the C oracle to ~1e-12 (different summation order / SVD), indices
exactly (absent exact distance ties).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def _transform(T, P):
    """open3d PointCloud::Transform -- cluster_icp.py:167"""
    H = np.hstack([P, np.ones((P.shape[0], 1))]) @ T.T
    return H[:, :3] / H[:, 3:4]


def _correspond(P, tree, n_t, r):
    """GetRegistrationResultAndCorrespondences (SearchHybrid radius=r, max_nn=1)"""
    n_s = P.shape[0]
    if n_t == 0 or n_s == 0:
        return np.full(n_s, -1, dtype=np.int64), 0.0, 0.0
    d, j = tree.query(P, k=1)
    d2 = d * d
    ok = d2 < r * r
    corr = np.where(ok, j, -1)
    c = int(ok.sum())
    if c == 0:
        return corr, 0.0, 0.0
    # rmse from exactly recomputed squared distances
    return corr, c / n_s, float(np.sqrt(d2[ok].sum() / c))


def _umeyama(A, B):
    """Eigen umeyama(src=A, dst=B, with_scaling=false) on (c,3) arrays"""
    mu_a, mu_b = A.mean(0), B.mean(0)
    sigma = (B - mu_b).T @ (A - mu_a) / A.shape[0]
    U, _, Vt = np.linalg.svd(sigma)
    D = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2] = -1
    R = U @ np.diag(D) @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mu_b - R @ mu_a
    return T


def icp_p2p(src, tgt, max_corr, init, max_iter=30, rel_fit=1e-6, rel_rmse=1e-6):
    """open3d 0.18.0 RegistrationICP, point-to-point (call site cluster_icp.py:157-159)."""
    if not max_corr > 0:
        raise RuntimeError("[Open3D Error] Invalid max_correspondence_distance.")
    src = np.asarray(src, dtype=np.float64).reshape(-1, 3)
    tgt = np.asarray(tgt, dtype=np.float64).reshape(-1, 3)
    T = np.array(init, dtype=np.float64).reshape(4, 4)
    P = _transform(T, src)
    tree = cKDTree(tgt) if tgt.shape[0] else None
    corr, fit, rmse = _correspond(P, tree, tgt.shape[0], max_corr)
    it = 0
    for i in range(max_iter):
        ok = corr >= 0
        U = _umeyama(P[ok], tgt[corr[ok]]) if ok.any() else np.eye(4)
        T = U @ T
        P = _transform(U, P)
        corr, fit2, rmse2 = _correspond(P, tree, tgt.shape[0], max_corr)
        stop = abs(fit - fit2) < rel_fit and abs(rmse - rmse2) < rel_rmse
        fit, rmse = fit2, rmse2
        it = i + 1
        if stop:
            break
    return dict(T=T, corr=corr.astype(np.int32), fitness=fit, rmse=rmse, iters=it, P=P)


def masked_icp(clusters_local, clusters_world, step_pc_np, matrices, visual=False, ori=False, scale=1.2, th=1,
               colors=None, _details=None):
    """cluster_icp.py:118-191 with the numpy lines kept as the reference wrote them."""
    world_clusters, new_matrices, details = [], [], []
    for c_local, c_world, matrix in zip(clusters_local, clusters_world, matrices):
        box = np.array([[np.min(c_world[:, 0]), np.max(c_world[:, 0])],
                        [np.min(c_world[:, 1]), np.max(c_world[:, 1])],
                        [np.min(c_world[:, 2]), np.max(c_world[:, 2])]])
        box_center = np.mean(box, axis=1)
        box_size = box[:, 1] - box[:, 0]
        box = np.vstack([box_center - 0.5 * scale * box_size, box_center + 0.5 * scale * box_size]).T
        mask = np.logical_and(step_pc_np[:, 0] > box[0, 0], step_pc_np[:, 0] < box[0, 1])
        mask = np.logical_and(mask, step_pc_np[:, 1] > box[1, 0])
        mask = np.logical_and(mask, step_pc_np[:, 1] < box[1, 1])
        mask = np.logical_and(mask, step_pc_np[:, 2] > box[2, 0])
        mask = np.logical_and(mask, step_pc_np[:, 2] < box[2, 1])
        masked_pc = step_pc_np[mask]
        r = icp_p2p(c_local, masked_pc, th, matrix, max_iter=10000)
        icp_matrix = r["T"].copy()
        if ori:
            icp_matrix[:3, 3] = np.asarray(matrix, dtype=np.float64)[:3, 3]
        world_clusters.append(_transform(icp_matrix, np.asarray(c_local, dtype=np.float64)))
        new_matrices.append(icp_matrix)
        r["mask_idx"] = np.nonzero(mask)[0]
        details.append(r)
    if _details is not None:
        _details["tiles"] = details
    return world_clusters, np.array(new_matrices)
