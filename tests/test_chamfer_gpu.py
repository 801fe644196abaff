"""GPU parity of the chamfer / 1-NN operator against the numpy oracle: nearest
indices bit-exact, distances bit-exact (same float32 operation order), loss and gradients to
float32 summation tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _clouds(seed, n1, n2, spread=0.3):
    rng = np.random.default_rng(seed)
    return rng.normal(0, spread, (n1, 3)).astype(np.float32), rng.normal(0, spread, (n2, 3)).astype(np.float32)


@pytest.mark.parametrize("norm", [1, 2])
@pytest.mark.parametrize("n1,n2", [(1, 1), (37, 5), (700, 1500), (5000, 5000)])
def test_knn1_bit_exact(norm, n1, n2):
    from autourdf_b200.chamfer import knn1
    from oracle import chamfer_oracle as O
    a, b = _clouds(n1 * 7 + n2, n1, n2)
    off = lambda n: torch.tensor([0, n], dtype=torch.int32, device="cuda")
    idx, dist = knn1(torch.from_numpy(a).cuda(), off(n1), torch.from_numpy(b).cuda(), off(n2), norm)
    oi, od = O.knn1(a, b, norm)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(dist.cpu().numpy(), od)


def test_knn1_groups_ties_and_empty():
    from autourdf_b200.chamfer import knn1
    from oracle import chamfer_oracle as O
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(1)
    q = [rng.normal(size=(n, 3)).astype(np.float32) for n in (300, 0, 40, 2000)]
    t = [rng.normal(size=(n, 3)).astype(np.float32) for n in (50, 10, 0, 3000)]
    t[0][7] = t[0][3]                                         # duplicate target -> lowest index must win
    q[0][0] = t[0][3]
    qo = np.concatenate([[0], np.cumsum([len(a) for a in q])]).astype(np.int32)
    to = np.concatenate([[0], np.cumsum([len(a) for a in t])]).astype(np.int32)
    idx, dist = knn1(torch.from_numpy(np.concatenate(q)).cuda(), torch.from_numpy(qo).cuda(),
                     torch.from_numpy(np.concatenate(t)).cuda(), torch.from_numpy(to).cuda(), 1)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for g in range(4):
        oi, od = O.knn1(q[g], t[g], 1)
        assert np.array_equal(idx[qo[g]:qo[g + 1]], oi) and np.array_equal(dist[qo[g]:qo[g + 1]], od)
    assert idx[0] == 3
    _, ii = cKDTree(t[3].astype(np.float64)).query(q[3].astype(np.float64), p=1)
    assert (ii == idx[qo[3]:qo[4]]).mean() > 0.999                   # float64 tree vs float32 scan: near-ties only


@pytest.mark.parametrize("norm", [1, 2])
def test_chamfer_loss_and_gradients(norm):
    from autourdf_b200.chamfer import chamfer_distance
    from oracle import chamfer_oracle as O
    rng = np.random.default_rng(3)
    x = rng.normal(0, 0.2, (2, 1800, 3)).astype(np.float32)
    y = rng.normal(0, 0.2, (2, 2300, 3)).astype(np.float32)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    yt = torch.from_numpy(y).cuda().requires_grad_(True)
    loss, normals = chamfer_distance(xt, yt, norm=norm)
    assert normals is None
    loss.backward()
    ol, ogx, ogy, _, _ = O.chamfer_distance(x, y, norm)
    assert abs(loss.item() - ol) <= 2e-6 * abs(ol)
    assert np.abs(xt.grad.cpu().numpy() - ogx).max() <= 1e-6 * max(1.0, np.abs(ogx).max()) + 1e-7
    assert np.abs(yt.grad.cpu().numpy() - ogy).max() <= 1e-6 * max(1.0, np.abs(ogy).max()) + 1e-7


def test_chamfer_reference_call_shape():
    """the call at mlp_reg.py:96: pred (1,M,3) with grad, y.unsqueeze(0) without"""
    from autourdf_b200.chamfer import chamfer_distance
    from oracle import chamfer_oracle as O
    from autourdf_b200 import synth
    b = synth.make_config("wx200", n_frames=2)
    pred = torch.from_numpy(b.box).cuda().unsqueeze(0).requires_grad_(True)          # predicted clusters, float32
    y = torch.from_numpy(b.tgt.astype(np.float32)).cuda()
    loss, _ = chamfer_distance(pred, y.unsqueeze(0), norm=1)
    loss.backward()
    ol, ogx, _, _, _ = O.chamfer_distance(b.box[None], b.tgt.astype(np.float32)[None], 1)
    assert abs(loss.item() - ol) <= 2e-6 * abs(ol)
    assert np.abs(pred.grad.cpu().numpy() - ogx).max() <= 1e-6
    with pytest.raises(ValueError):
        chamfer_distance(pred, y.unsqueeze(0), norm=3)


@pytest.mark.parametrize("N,P1,P2", [(1, 1, 1), (1, 37, 5), (3, 700, 1500), (1, 5000, 5000), (2, 513, 64), (1, 20000, 3000)])
def test_fused_chamfer_matches_packed_knn_and_is_repeatable(N, P1, P2):
    """the one-launch forward: nearest-neighbour indices (kept for the backward) identical to the packed
    aurdf_nn_f32 path (itself bit-exact against the oracle), loss within float32 summation error of the oracle,
    and -- the arrival counters reset themselves -- bit-identical on repeated calls"""
    from autourdf_b200 import chamfer as ch
    from oracle import chamfer_oracle as O
    rng = np.random.default_rng(N * 1000 + P1 + P2)
    x = rng.normal(0, 0.2, (N, P1, 3)).astype(np.float32)
    y = rng.normal(0, 0.2, (N, P2, 3)).astype(np.float32)
    y[0, 0] = x[0, 0]                                                # one exact coincidence
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    losses = []
    for norm in (1, 2):
        ol = O.chamfer_distance(x, y, norm)[0]
        for rep in range(3):
            xr = xt.clone().requires_grad_(True)
            loss, _ = ch.chamfer_distance(xr, yt, norm=norm)
            loss.backward()
            losses.append((norm, loss.item(), xr.grad.abs().sum().item()))
            assert abs(loss.item() - ol) <= 4e-6 * abs(ol) + 1e-12
        assert losses[-1][1] == losses[-2][1] == losses[-3][1], "loss differs between identical calls"
        xo = torch.arange(0, (N + 1) * P1, P1, dtype=torch.int32, device="cuda")
        yo = torch.arange(0, (N + 1) * P2, P2, dtype=torch.int32, device="cuda")
        ix, dx = ch.knn1(xt.reshape(-1, 3), xo, yt.reshape(-1, 3), yo, norm)
        iy, dy = ch.knn1(yt.reshape(-1, 3), yo, xt.reshape(-1, 3), xo, norm)
        # indices saved by the fused forward (through the autograd node of a fresh call)
        xr = xt.clone().requires_grad_(True)
        loss, _ = ch.chamfer_distance(xr, yt, norm=norm)
        idx = loss.grad_fn.saved_tensors[2]
        assert torch.equal(idx[:N * P1], ix) and torch.equal(idx[N * P1:], iy)
        for red in (("sum", "mean"), ("mean", "sum"), ("sum", "sum")):
            l2, _ = ch.chamfer_distance(xt, yt, norm=norm, batch_reduction=red[0], point_reduction=red[1])
            cx = dx.reshape(N, P1).double().sum(1) / (P1 if red[1] == "mean" else 1)
            cy = dy.reshape(N, P2).double().sum(1) / (P2 if red[1] == "mean" else 1)
            want = (cx.sum() + cy.sum()) / (N if red[0] == "mean" else 1)
            assert abs(l2.item() - want.item()) <= 4e-6 * abs(want.item()) + 1e-12
