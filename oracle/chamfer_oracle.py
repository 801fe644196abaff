"""numpy restatement of pytorch3d 0.7.7 ``loss.chamfer_distance(x, y, norm=1|2)`` with default
reductions, and of the ``ops.knn_points(K=1)`` forward/backward under it (call sites: reference
PointCloud/mlp_reg.py:96, Sim/evaluation.py:81).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: pytorch3d is not vendored in the reference and not
installable here; this follows its published algorithm (csrc/knn/knn_cpu.cpp: float32 accumulation
``dist += |diff|`` / ``diff*diff`` over d = 0,1,2, strict '<' so the first minimum wins; backward
``sign(diff) * grad`` / ``2 diff grad`` to p1 and its negative scattered to p2) and is cross-checked
against scipy cKDTree (Minkowski p=1 / p=2) in tests.
"""
import numpy as np


def knn1(p1, p2, norm=1):
    """p1 (n1,3), p2 (n2,3) float32 -> idx (n1,) int32, dist (n1,) float32"""
    p1 = np.asarray(p1, dtype=np.float32); p2 = np.asarray(p2, dtype=np.float32)
    if p2.shape[0] == 0:
        return np.full(p1.shape[0], -1, np.int32), np.full(p1.shape[0], np.inf, np.float32)
    idx = np.empty(p1.shape[0], np.int32); dist = np.empty(p1.shape[0], np.float32)
    step = max(1, (1 << 22) // max(p2.shape[0], 1))
    for s in range(0, p1.shape[0], step):
        diff = p1[s:s + step, None, :] - p2[None, :, :]                      # float32
        e = np.abs(diff) if norm == 1 else diff * diff
        d = (e[..., 0] + e[..., 1]) + e[..., 2]                              # float32, d = 0,1,2 order
        j = d.argmin(1)                                                      # first minimum
        idx[s:s + step] = j
        dist[s:s + step] = d[np.arange(d.shape[0]), j]
    return idx, dist


def chamfer_distance(x, y, norm=1):
    """x (N,P1,3), y (N,P2,3) -> (loss float64 of float32 terms, grad_x, grad_y, idx_x, idx_y)"""
    x = np.asarray(x, dtype=np.float32); y = np.asarray(y, dtype=np.float32)
    N, P1, P2 = x.shape[0], x.shape[1], y.shape[1]
    loss = 0.0
    gx = np.zeros_like(x, dtype=np.float64); gy = np.zeros_like(y, dtype=np.float64)
    ix_all, iy_all = [], []
    for n in range(N):
        ix, dx = knn1(x[n], y[n], norm)
        iy, dy = knn1(y[n], x[n], norm)
        loss += dx.astype(np.float64).sum() / P1 / N + dy.astype(np.float64).sum() / P2 / N
        dfx = (x[n] - y[n][ix]).astype(np.float64)
        dfy = (y[n] - x[n][iy]).astype(np.float64)
        g1 = (np.sign(dfx) if norm == 1 else 2 * dfx) / (P1 * N)
        g2 = (np.sign(dfy) if norm == 1 else 2 * dfy) / (P2 * N)
        gx[n] += g1
        np.add.at(gy[n], ix, -g1)
        gy[n] += g2
        np.add.at(gx[n], iy, -g2)
        ix_all.append(ix); iy_all.append(iy)
    return loss, gx, gy, np.stack(ix_all), np.stack(iy_all)
