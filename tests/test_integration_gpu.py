"""In-situ test of the import swaps of INTEGRATION.md section 1.

The reference's registration step -- ``train()`` (AutoURDF PointCloud/mlp_reg.py:17-119: pose MLP ->
rotation conversion -> ``calculate_pc`` -> ``chamfer_distance(norm=1)`` -> ``loss.backward()`` -> Adam) followed by
the ``--mlp_icp`` frame step (``match()``, :322-328: ``masked_icp`` -> ``resample_cluster``) -- is restated here
line for line around the operator table it imports, and run twice on the GPU with the same seeds:

  * with the engine's operators swapped in at exactly the import sites INTEGRATION.md lists
    (``autourdf_b200.dq_func`` / ``mlp_reg`` / ``chamfer`` / ``cluster_icp``), and
  * with the restated reference operators (plain torch expressions of pytorch3d / dq_func / calculate_pc, a
    cdist-based chamfer, the CPU oracle for ICP and k-means).

The loop itself cannot be imported at run time: /root/reference does not exist on the GPU box and open3d /
pytorch3d are not installable (tests/golden/make_golden.py runs the reference's own functions in the build
container; this test is its run-time counterpart for the training loop).  What it checks is what the swap could
break: gradients reach the MLP through every swapped operator (both ``--r q`` and ``--r dq``), the loss follows
the reference operators' loss epoch by epoch and decreases, and the ICP step returns the oracle's poses.
The pose MLPs (model_utils.py) are outside the hot path; two small residual MLPs stand in for them.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _QReg(torch.nn.Module):
    """stand-in for model_utils.QRegMLP: (N, 7) translation + quaternion -> (translation, quaternion)"""

    def __init__(self, hidden=64):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(7, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, 7))
        with torch.no_grad():
            self.net[2].weight.mul_(0.01)
            self.net[2].bias.zero_()

    def forward(self, x):
        o = x + self.net(x)
        return o[:, :3], o[:, 3:]


class _DQReg(torch.nn.Module):
    """stand-in for model_utils.DQRegMLP: (N, 8) dual quaternion -> (N, 8)"""

    def __init__(self, hidden=64):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(8, hidden), torch.nn.Tanh(), torch.nn.Linear(hidden, 8))
        with torch.no_grad():
            self.net[2].weight.mul_(0.01)
            self.net[2].bias.zero_()

    def forward(self, x):
        return x + self.net(x)


def _ref_ops():
    """the reference's operators, restated in torch (oracle/pt3d_torch.py, mlp_reg.py:155-170, chamfer via cdist)"""
    from oracle import pt3d_torch as P

    def calculate_pc(local_clusters, matrices):                       # mlp_reg.py:155-170
        return [ic @ matrices[i][:3, :3].T + matrices[i][:3, 3] for i, ic in enumerate(local_clusters)]

    def chamfer_distance(x, y, norm=1):                               # pytorch3d semantics, point / batch mean
        d = torch.cdist(x, y, p=norm)
        if norm == 2:
            d = d * d
        return d.min(2).values.mean(1).mean() + d.min(1).values.mean(1).mean(), None

    return dict(matrix_to_quaternion=P.matrix_to_quaternion, quaternion_to_matrix=P.quaternion_to_matrix,
                transform_to_dualquat=P.transform_to_dualquat, dualquat_to_transform=P.dualquat_to_transform,
                calculate_pc=calculate_pc, chamfer_distance=chamfer_distance)


def _engine_ops():
    """the same names from the engine"""
    from autourdf_b200 import dq_func as D
    from autourdf_b200.chamfer import chamfer_distance
    from autourdf_b200.mlp_reg import calculate_pc
    return dict(matrix_to_quaternion=D.matrix_to_quaternion, quaternion_to_matrix=D.quaternion_to_matrix,
                transform_to_dualquat=D.transform_to_dualquat, dualquat_to_transform=D.dualquat_to_transform,
                calculate_pc=calculate_pc, chamfer_distance=chamfer_distance)


def _train(ops, rot, m, y, model, clusters, epochs, lr=2e-4):
    """mlp_reg.py:17-119 (`train`), early stopping / scheduler left out (they only read loss.item())"""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses, best, best_m, best_pcd = [], 1000.0, None, None
    for _ in range(epochs):
        m2 = m.clone()
        if rot == "q":                                                # :60-66
            q = ops["matrix_to_quaternion"](m2[:, :3, :3])
            t, r = model(torch.cat([m2[:, :3, 3], q], dim=1))
            m2[:, :3, :3] = ops["quaternion_to_matrix"](r)
            m2[:, :3, 3] = t
        else:                                                         # :78-84
            m2 = ops["dualquat_to_transform"](model(ops["transform_to_dualquat"](m2)))
        pred_list = ops["calculate_pc"](clusters, m2)                 # :93
        pred = torch.cat(pred_list, dim=0).unsqueeze(0)
        loss, _ = ops["chamfer_distance"](pred, y.unsqueeze(0), norm=1)   # :96
        losses.append(loss.item())
        if loss.item() < best:
            best, best_m, best_pcd = loss.item(), m2, pred_list
        opt.zero_grad()
        loss.backward()                                               # :114-116
        opt.step()
    return losses, best_m.detach(), [p.detach().cpu().numpy() for p in best_pcd]


@pytest.mark.parametrize("rot", ["q", "dq"])
def test_train_loop_with_engine_operators_follows_reference_operators(rot):
    from autourdf_b200 import synth
    b = synth.make_config("wx200", n_frames=3)
    K = b.n_clusters
    dev = torch.device("cuda")
    clusters = [torch.tensor(b.src[b.src_off[k]:b.src_off[k + 1]], dtype=torch.float32, device=dev) for k in range(K)]
    m = torch.tensor(b.init_T[:K], dtype=torch.float32, device=dev)                    # step matrices (mlp_reg.py:286)
    y = torch.tensor(b.tgt[b.tgt_off[0]:b.tgt_off[1]], dtype=torch.float32, device=dev)  # target cloud (:296)
    runs = {}
    for name, ops in (("engine", _engine_ops()), ("reference", _ref_ops())):
        torch.manual_seed(0)
        model = (_QReg() if rot == "q" else _DQReg()).to(dev)
        runs[name] = _train(ops, rot, m, y, model, clusters, epochs=40, lr=3e-4)
        grads = [p.grad for p in model.parameters()]
        assert all(g is not None and torch.isfinite(g).all() for g in grads), f"{name}: no gradient reached the MLP"
        assert any(g.abs().max().item() > 0 for g in grads)
    le, lr_ = np.array(runs["engine"][0]), np.array(runs["reference"][0])
    assert le[-1] < 0.99 * le[0], f"loss did not decrease with the engine's operators: {le[0]} -> {le[-1]}"
    # the same optimisation: float32 operators, so the trajectories agree closely at first and stay close
    assert np.abs(le[:5] - lr_[:5]).max() <= 1e-4 * lr_[0]
    assert np.abs(le - lr_).max() <= 2e-2 * lr_[0]
    assert (runs["engine"][1] - runs["reference"][1]).abs().max().item() <= 1e-2


def test_mlp_icp_frame_step_matches_oracle(oracle):
    """match() with --mlp_icp (mlp_reg.py:322-328): train() -> masked_icp -> resample_cluster, engine vs oracle"""
    from autourdf_b200 import synth
    from autourdf_b200.cluster_icp import masked_icp
    from autourdf_b200.mlp_reg import resample_cluster
    from oracle import kmeans_oracle as KO
    b = synth.make_config("wx200", n_frames=3)
    K = b.n_clusters
    dev = torch.device("cuda")
    step_cluster_np = [b.src[b.src_off[k]:b.src_off[k + 1]] for k in range(K)]
    clusters = [torch.tensor(c, dtype=torch.float32, device=dev) for c in step_cluster_np]
    m = torch.tensor(b.init_T[:K], dtype=torch.float32, device=dev)
    target_pcd_np = b.tgt[b.tgt_off[0]:b.tgt_off[1]]
    y = torch.tensor(target_pcd_np, dtype=torch.float32, device=dev)
    torch.manual_seed(1)
    _, step_m, pred_pcd_np = _train(_engine_ops(), "q", m, y, _QReg().to(dev), clusters, epochs=15, lr=1e-3)
    step_m_np = step_m.cpu().numpy()                                  # float32 (K,4,4), :322
    world, matrices = masked_icp(step_cluster_np, pred_pcd_np, target_pcd_np, step_m_np, False, ori=False)   # :325
    wo, mo = oracle.masked_icp(step_cluster_np, pred_pcd_np, target_pcd_np, step_m_np, False, ori=False)
    assert matrices.dtype == np.float64 and matrices.shape == (K, 4, 4)
    assert np.abs(matrices - mo).max() <= 1e-5
    for a, c in zip(world, wo):
        assert a.shape == c.shape and (a.shape[0] == 0 or np.abs(a - c).max() <= 1e-5)
    new_seg = resample_cluster(target_pcd_np, 0, K, matrices)        # :326
    ref_seg, _ = KO.resample_cluster(target_pcd_np, K, mo)
    assert [s.shape for s in new_seg] == [s.shape for s in ref_seg]
    for a, c in zip(new_seg, ref_seg):
        assert a.shape[0] == 0 or np.abs(a - c).max() <= 1e-5
