"""numpy restatement of ``resample_cluster`` (reference PointCloud/mlp_reg.py:172-237, normal=False):
``sklearn.cluster.k_means(pc_np, init=xyz, n_clusters=K, n_init=1)`` seeded with the fitted cluster
origins (:204), then every cluster moved into its local frame with ``inv(matrix)`` (:211-213).

TEST INFRASTRUCTURE ONLY.  PINNED: scikit-learn is installed in the build container, so this
restatement of its Lloyd iteration (centre the data, E-step argmin of |c|^2 - 2 x.c with the first
minimum winning, M-step means, empty clusters re-seeded with the farthest points, stop on unchanged
labels or on sum of squared centre shifts <= 1e-4 * mean feature variance, final E-step unless the
labels had converged) is checked against ``sklearn.cluster.k_means`` itself in tests/test_oracle_cpu.py
(labels identical, centres to 1e-12).  Reference pin is scikit-learn==1.5.2 (requirements.txt:10);
the container has 1.9.0 -- the Lloyd loop is unchanged between them.
"""
import numpy as np


def k_means_lloyd(X, init, max_iter=300, tol=1e-4):
    """-> (centers (K,3), labels (N,) int32, inertia, n_iter)"""
    X = np.asarray(X, dtype=np.float64)
    K = init.shape[0]
    mean = X.mean(axis=0)
    Xc = X - mean
    C = np.asarray(init, dtype=np.float64) - mean
    tol_ = np.mean(np.var(Xc, axis=0)) * tol
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for it in range(max_iter):
        d = (C * C).sum(1)[None, :] - 2.0 * (Xc @ C.T)
        labels = d.argmin(1).astype(np.int32)
        w = np.bincount(labels, minlength=K).astype(np.float64)
        Cn = np.zeros_like(C)
        np.add.at(Cn, labels, Xc)
        empty = np.nonzero(w == 0)[0]
        if empty.size:                                   # _relocate_empty_clusters_dense
            dist = ((Xc - C[labels]) ** 2).sum(1)
            far = np.argsort(-dist, kind="stable")[:empty.size]
            for e, fi in zip(empty, far):
                old = labels[fi]
                Cn[old] -= Xc[fi]
                Cn[e] = Xc[fi]
                w[e] = 1.0
                w[old] -= 1.0
        Cn /= w[:, None]
        shift_tot = ((Cn - C) ** 2).sum()
        C = Cn
        n_iter = it + 1
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift_tot <= tol_:
            break
        labels_old = labels
    if not strict:
        d = (C * C).sum(1)[None, :] - 2.0 * (Xc @ C.T)
        labels = d.argmin(1).astype(np.int32)
    inertia = float(((Xc - C[labels]) ** 2).sum())
    return C + mean, labels, inertia, n_iter


def resample_cluster(pc_np, n_clusters, matrices):
    """mlp_reg.py:172-237 on arrays: -> list of (n_k,3) float64 clusters in their local frames"""
    pc_np = np.asarray(pc_np, dtype=np.float64)
    matrices = np.asarray(matrices, dtype=np.float64)
    xyz = matrices[:, :3, 3]
    _, labels, _, _ = k_means_lloyd(pc_np, xyz[:n_clusters])
    out = []
    for i in range(n_clusters):
        pts = pc_np[labels == i]
        inv = np.linalg.inv(matrices[i])
        out.append((inv @ np.hstack([pts, np.ones((pts.shape[0], 1))]).T)[:3].T)
    return out, labels
