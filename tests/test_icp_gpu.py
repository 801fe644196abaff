"""GPU parity: CUDA cluster-ICP sweep (through the C ABI) vs the CPU oracle and the golden
vectors produced by the reference's own cluster_icp.py.  Gates (north_star / SURVEY 8d):
correspondence indices bit-exact, iteration counts equal, poses within 1e-5."""
import os

import numpy as np
import pytest

from helpers import assert_parity, cuda_sweep, oracle_sweep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    from autourdf_b200 import synth
    return synth


def test_wx200_c1_parity(oracle, synth):
    b = synth.make_config("wx200")
    assert_parity(cuda_sweep(b), oracle_sweep(oracle, b), what="wx200", batch=b)


def test_wx200_5_c2_parity(oracle, synth):
    b = synth.make_config("wx200_5")
    g, o = cuda_sweep(b), oracle_sweep(oracle, b, use_kdtree=True)
    assert_parity(g, o, what="wx200_5", batch=b)
    assert (o["ntgt"] == 0).any(), "config is expected to contain an empty-mask tile"


def test_franka_c3_parity(oracle, synth):
    b = synth.make_config("franka")
    assert_parity(cuda_sweep(b), oracle_sweep(oracle, b, use_kdtree=True), what="franka", batch=b)


def test_allegro_c4_parity(oracle, synth):
    """the full C4 config (5 sequences, 2160 tiles), rank-deficient tiles included"""
    b = synth.make_config("allegro_hand")
    o = oracle_sweep(oracle, b, use_kdtree=True)
    assert_parity(cuda_sweep(b), o, what="allegro", batch=b)
    assert (o["cond"] <= 1e-6).sum() >= 10, "config is expected to contain rank-deficient tiles"


@pytest.mark.parametrize("n_points,n_clusters,n_frames", [(1024, 8, 11), (4096, 32, 11), (16384, 32, 6), (16384, 128, 6),
                                                          (65536, 128, 6), (65536, 8, 4)])
def test_c5_sweep_points_parity(oracle, synth, n_points, n_clusters, n_frames):
    """points of the C5 sweep (BASELINE.json configs[4]): small tiles, 512-point tiles (general kernel),
    8192-point tiles (thread-block-cluster variant)"""
    oracle.use_all_host_threads()
    b = synth.make_batch(n_points=n_points, n_clusters=n_clusters, n_seq=1, n_frames=n_frames, dof=5, cid=5)
    assert_parity(cuda_sweep(b), oracle_sweep(oracle, b, use_kdtree=True), what=f"C5 {n_points}x{n_clusters}", batch=b)


def test_f32_storage_same_result(oracle, synth):
    """inputs are float32-representable, so float32 storage must give identical results"""
    import torch
    b = synth.make_config("wx200")
    src32 = b.src.astype(np.float32).astype(np.float64)
    b.src = src32                       # make the local clusters float32-representable too
    o = oracle_sweep(oracle, b)
    assert_parity(cuda_sweep(b, pts_dtype=torch.float32), o, what="f32 storage")
    assert_parity(cuda_sweep(b, pts_dtype=torch.float64), o, what="f64 storage")


@pytest.mark.parametrize("ori", [False, True])
def test_golden_masked_icp(golden_dir, ori):
    """the drop-in masked_icp vs outputs of the reference's own masked_icp (make_golden.py)"""
    from autourdf_b200.cluster_icp import masked_icp
    z = np.load(os.path.join(golden_dir, "masked_icp.npz"))
    K = int(z["K"])
    F = z["tgt_off"].shape[0] - 1
    for f in range(F):
        tiles = range(f * K, (f + 1) * K)
        cl = [z["src"][z["src_off"][t]:z["src_off"][t + 1]] for t in tiles]
        cw = [z["box"][z["box_off"][t]:z["box_off"][t + 1]] for t in tiles]
        cloud = z["tgt"][z["tgt_off"][f]:z["tgt_off"][f + 1]]
        mats = z["init_T"][f * K:(f + 1) * K].astype(np.float32)
        w, m = masked_icp(cl, cw, cloud, mats, False, ori=ori)
        assert m.dtype == np.float64 and m.shape == (K, 4, 4)
        assert np.abs(m - z[f"f{f}_ori{int(ori)}_T"]).max() <= 1e-5
        assert np.abs(np.concatenate(w) - z[f"f{f}_ori{int(ori)}_world"]).max() <= 1e-5
        assert [x.shape for x in w] == [c.shape for c in cl]


def test_golden_masked_icp_f64_box_scale_th(golden_dir):
    from autourdf_b200.cluster_icp import masked_icp
    z = np.load(os.path.join(golden_dir, "masked_icp.npz"))
    K = int(z["K"])
    for f in range(z["tgt_off"].shape[0] - 1):
        tiles = range(f * K, (f + 1) * K)
        cl = [z["src"][z["src_off"][t]:z["src_off"][t + 1]] for t in tiles]
        cw = [z["box"][z["box_off"][t]:z["box_off"][t + 1]].astype(np.float64) for t in tiles]
        cloud = z["tgt"][z["tgt_off"][f]:z["tgt_off"][f + 1]]
        mats = z["init_T"][f * K:(f + 1) * K].astype(np.float32)
        w, m = masked_icp(cl, cw, cloud, mats, False, ori=False, scale=1.5, th=0.02)
        assert np.abs(m - z[f"f{f}_f64box_T"]).max() <= 1e-5
        assert np.abs(np.concatenate(w) - z[f"f{f}_f64box_world"]).max() <= 1e-5


def test_edge_cases(oracle):
    """empty source cluster, empty mask, single point, threshold rejection, zip truncation"""
    from autourdf_b200.cluster_icp import masked_icp
    rng = np.random.default_rng(3)
    cloud = rng.uniform(-0.2, 0.2, size=(500, 3)).astype(np.float32).astype(np.float64)
    I = np.eye(4, dtype=np.float32)
    far = np.eye(4, dtype=np.float32); far[:3, 3] = 5.0
    cl = [cloud[:40] * 0.5, np.zeros((0, 3)), cloud[40:41].copy(), cloud[50:120] * 0.9, cloud[:30].copy()]
    mats = [I, I, I, I, far]
    cw = [(c @ m[:3, :3].T + m[:3, 3]).astype(np.float32) if c.shape[0] else np.zeros((1, 3), np.float32)
          for c, m in zip(cl, mats)]
    for th in (1, 0.004):
        dg, do = {}, {}
        wg, mg = masked_icp(cl, cw, cloud, mats + [I, I], th=th, _details=dg)   # extra matrices: truncated
        wo, mo = oracle.masked_icp(cl, cw, cloud, mats, th=th, _details=do)
        assert len(wg) == 5 and mg.shape == (5, 4, 4)
        assert np.array_equal(dg["corr"], do["corr"]) and np.array_equal(dg["iters"], do["iters"])
        assert np.array_equal(dg["ntgt"], do["ntgt"])
        assert np.abs(mg - mo).max() <= 1e-5
        assert np.array_equal(mg[4], far.astype(np.float64))            # empty mask -> init unchanged
        assert wg[1].shape == (0, 3)
    with pytest.raises(ValueError):
        masked_icp([cl[0]], [np.zeros((0, 3), np.float32)], cloud, [I])
    with pytest.raises(RuntimeError):
        masked_icp([cl[0]], [cw[0]], cloud, [I], th=0)


def test_registration_icp_callsites(oracle):
    """link.py:113 (th=1, init=I, 1e5 iters) and evaluation.py:358 (th=0.01, 2e4 iters)"""
    from autourdf_b200.cluster_icp import registration_icp
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(11)
    tgt = rng.uniform(-0.1, 0.1, size=(3000, 3))
    Rm = Rotation.from_rotvec([0.02, -0.03, 0.025]).as_matrix()
    src = (tgt[:2500] - [0.002, 0.001, -0.0015]) @ Rm + rng.normal(0, 2e-4, size=(2500, 3))
    for th, it in ((1, 100000), (0.01, 20000)):
        g = registration_icp(src, tgt, th, np.eye(4), max_iteration=it)
        o = oracle.icp_p2p(src, tgt, th, np.eye(4), max_iter=it, use_kdtree=True)
        assert g.iterations == o["iters"]
        i = np.nonzero(o["corr"] >= 0)[0]
        assert np.array_equal(g.correspondence_set, np.stack([i, o["corr"][i]], 1))
        assert np.abs(g.transformation - o["T"]).max() <= 1e-5
        assert abs(g.fitness - o["fitness"]) <= 1e-12 and abs(g.inlier_rmse - o["rmse"]) <= 1e-9


def test_large_tiles_streaming_and_spill(oracle, synth):
    """tiles whose target exceeds one shared-memory chunk (streamed via TMA) and whose source
    exceeds the shared-memory bound (workspace spill) -- a C5 sweep point, small frame count"""
    b = synth.make_batch(n_points=16384, n_clusters=4, n_seq=1, n_frames=3, dof=5, cid=5)
    assert np.diff(b.src_off).max() > 2048
    assert_parity(cuda_sweep(b), oracle_sweep(oracle, b, use_kdtree=True), what="large tiles", batch=b)


def test_permutation_invariance(synth):
    """relabelling the target points must only relabel the correspondences"""
    b = synth.make_config("wx200", n_frames=3)
    g0 = cuda_sweep(b)
    rng = np.random.default_rng(0)
    b2 = synth.make_config("wx200", n_frames=3)
    perms = []
    for f in range(b.n_frames):
        s, e = b.tgt_off[f], b.tgt_off[f + 1]
        p = rng.permutation(e - s)
        b2.tgt[s:e] = b.tgt[s:e][p]
        perms.append(p)
    g1 = cuda_sweep(b2)
    for t in range(b.n_tiles):
        f = b.tile_frame[t]
        c0 = g0["corr"][b.src_off[t]:b.src_off[t + 1]]
        c1 = g1["corr"][b.src_off[t]:b.src_off[t + 1]]
        ok = c0 >= 0
        assert np.array_equal(c0 >= 0, c1 >= 0)
        assert np.array_equal(perms[f][c1[ok]], c0[ok])
    assert np.abs(g0["T"] - g1["T"]).max() <= 1e-9


def _shard_worker(rank, world, port, q):
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)   # 2 ranks share the one GPU of the test box
    try:
        from autourdf_b200 import synth
        from autourdf_b200.dist import cuda_run_local, sharded_sweep
        torch.cuda.set_device(0)
        b = synth.make_config("wx200")
        run = cuda_run_local(torch.device("cuda", 0))
        allr, local, _ = sharded_sweep(b, run, device="cpu")
        full = run(b)
        ok = (np.array_equal(allr["T"], full["T"].cpu().numpy()) and
              np.array_equal(allr["iters"], full["iters"].cpu().numpy()))
        q.put((rank, bool(ok), ""))
    except Exception as e:
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_two_ranks_bit_identical_to_single():
    """the sharded result must be bit-identical to the 1-GPU result"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res


def test_single_gpu_sharded_entry_points(synth):
    """world == 1 through the documented CUDA entry points: sharded_sweep(b, cuda_run_local()) and the
    device-resident ShardedSweep plan give the plain sweep's poses bit for bit"""
    import torch
    from autourdf_b200.dist import ShardedSweep, cuda_run_local, sharded_sweep
    b = synth.make_config("wx200", n_frames=4)
    g = cuda_sweep(b)
    allr, local, span = sharded_sweep(b, cuda_run_local())
    assert span == (0, b.n_frames) and np.array_equal(allr["T"], g["T"]) and np.array_equal(allr["iters"], g["iters"])
    for by in ("frames", "tiles"):
        sh = ShardedSweep(b, torch.device("cuda"), by=by)
        sh.run()
        torch.cuda.synchronize()
        assert np.array_equal(sh.poses(), g["T"])


def test_host_path_frame_blocks_match_device_path(synth):
    """aurdf_icp_sweep_host cuts a large batch into frame blocks on separate streams (copies of one
    block overlap the kernels of another); tiles are independent, so every output must be
    bit-identical to the single-launch device path -- pageable and pinned caller buffers alike."""
    import torch
    from autourdf_b200 import cluster_icp as ci
    b = synth.make_config("wx200_5")
    g = cuda_sweep(b)
    host = ci.HostSweep()
    args = (b.src, b.src_off, b.tgt, b.tgt_off, b.tile_frame, b.box, b.box_off, b.init_T)
    for rep in range(2):          # second call: capacity hints of the first are reused
        o = host.run(*args)
        for k in ("T", "world", "corr", "fitness", "rmse", "iters", "ntgt"):
            assert np.array_equal(o[k], g[k]), f"host path (pageable, call {rep}) differs in {k}"
    keep = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in args]
    o = host.run(*[t.numpy() for t in keep])
    for k in ("T", "world", "corr", "fitness", "rmse", "iters", "ntgt"):
        assert np.array_equal(o[k], g[k]), f"host path (pinned) differs in {k}"
    h2d, d2h = host.copy_bytes()
    assert h2d > 0 and d2h > 0


def test_tile_class_boundaries(oracle):
    """Tiles at the seams between the small-tile kernel (n_s <= 320, n_t <= 760; float64 targets in
    shared memory up to 384) and the general kernel, plus the shapes that change the small kernel's
    lane split (tiny n_s), its padding (odd / even / single targets) and its second round
    (n_s > 128): every one must give the oracle's correspondences, iteration counts and poses."""
    import torch
    from scipy.spatial.transform import Rotation
    from autourdf_b200 import cluster_icp as ci
    cases = [(5, 3), (1, 1), (3, 2), (17, 2), (2, 700), (320, 760), (321, 760), (320, 761), (100, 385), (100, 384),
             (64, 500), (300, 41), (128, 129), (129, 128), (257, 300), (40, 759), (7, 64), (33, 33)]
    rng = np.random.default_rng(5)
    srcs, tgts, inits = [], [], []
    for ns, nt in cases:
        tgt = rng.uniform(-0.1, 0.1, size=(nt, 3)).astype(np.float32).astype(np.float64)
        R = Rotation.from_rotvec(rng.normal(scale=0.02, size=3)).as_matrix()
        pick = tgt[rng.integers(0, nt, size=ns)] + rng.normal(scale=2e-3, size=(ns, 3))
        srcs.append(((pick - rng.normal(scale=2e-3, size=3)) @ R).astype(np.float32).astype(np.float64))
        tgts.append(tgt)
        T0 = np.eye(4)
        T0[:3, 3] = rng.normal(scale=1e-3, size=3)
        inits.append(T0)
    dev = torch.device("cuda")

    def run(idx, max_iter):
        s_off = np.zeros(len(idx) + 1, np.int32)
        t_off = np.zeros(len(idx) + 1, np.int32)
        s_off[1:] = np.cumsum([srcs[i].shape[0] for i in idx])
        t_off[1:] = np.cumsum([tgts[i].shape[0] for i in idx])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        r = ci.icp_sweep(t(np.concatenate([srcs[i] for i in idx])), t(s_off), t(np.concatenate([tgts[i] for i in idx])),
                         t(t_off), t(np.arange(len(idx), dtype=np.int32)), None, None,
                         t(np.stack([inits[i] for i in idx])), max_src_per_tile=int(np.diff(s_off).max()),
                         max_iter=max_iter)
        torch.cuda.synchronize()
        return r, s_off

    # the initial correspondence pass of every shape (max_iter = 0: no pose fit, so the 2-3 point
    # tiles, whose Kabsch problem is rank-deficient, can be compared too)
    r, s_off = run(list(range(len(cases))), 0)
    corr, ntg = r.corr.cpu().numpy(), r.ntgt.cpu().numpy()
    for n, i in enumerate(range(len(cases))):
        o = oracle.icp_p2p(srcs[i], tgts[i], 1.0, inits[i], max_iter=0)
        assert ntg[n] == cases[i][1]
        assert np.array_equal(corr[s_off[n]:s_off[n + 1]], o["corr"]), f"initial pass differs for (n_s, n_t) = {cases[i]}"
    # full ICP on the well-posed shapes
    well = [i for i, (ns, nt) in enumerate(cases) if ns >= 7 and nt >= 33]
    r, s_off = run(well, 40)
    corr, iters, T = r.corr.cpu().numpy(), r.iters.cpu().numpy(), r.T.cpu().numpy()
    fit, rmse = r.fitness.cpu().numpy(), r.rmse.cpu().numpy()
    for n, i in enumerate(well):
        o = oracle.icp_p2p(srcs[i], tgts[i], 1.0, inits[i], max_iter=40)
        assert iters[n] == o["iters"], f"iterations differ for {cases[i]}: {iters[n]} vs {o['iters']}"
        assert np.array_equal(corr[s_off[n]:s_off[n + 1]], o["corr"]), f"correspondences differ for {cases[i]}"
        assert np.abs(T[n] - o["T"]).max() <= 1e-5
        assert abs(fit[n] - o["fitness"]) <= 1e-12 and abs(rmse[n] - o["rmse"]) <= 1e-9


def test_grid_search_adversarial(oracle):
    """Shapes that stress the grid-pruned exact search of icp_grid_kernel (medium / large tiles): sources far
    outside the targets' bounding box (every block widening path), flat and rod-like target sets (degenerate
    grid axes), duplicated targets (exact ties -> lowest index), a lattice (many near-ties), a rank-deficient
    large tile (strict pose fit in the one-CTA variant), a tight correspondence threshold, and the same shapes
    both as one-CTA tiles and -- with a large tile in the batch -- as 8-CTA cluster tiles."""
    import torch
    from scipy.spatial.transform import Rotation
    from autourdf_b200 import cluster_icp as ci
    rng = np.random.default_rng(17)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)

    def moved(tgt, ns, rot=0.03, shift=2e-3, noise=5e-4):
        R = Rotation.from_rotvec(rng.normal(scale=rot, size=3)).as_matrix()
        pick = tgt[rng.integers(0, tgt.shape[0], size=ns)] + rng.normal(scale=noise, size=(ns, 3))
        return f32((pick - rng.normal(scale=shift, size=3)) @ R)

    cases = {}
    surf = rng.uniform(-0.1, 0.1, size=(3000, 3)); surf[:, 2] = 0.02 * np.sin(20 * surf[:, 0])     # a curved sheet
    cases["sheet"] = (moved(f32(surf), 900), f32(surf), 1.0)
    vol = f32(rng.uniform(-0.05, 0.05, size=(2500, 3)))
    cases["outside"] = (f32(moved(vol, 700) + [0.03, -0.02, 0.025]), vol, 1.0)                     # starts half a box away
    far = (f32(moved(vol, 700) + [0.4, -0.3, 0.25]), vol)   # 0.5 m away: every source sees the same corner (rank-deficient
                                                            # fit), so only the initial correspondence pass is compared
    flat = rng.uniform(-0.1, 0.1, size=(1500, 3)); flat[:, 1] = rng.normal(scale=1e-4, size=1500)   # one grid cell thick
    cases["flat"] = (moved(f32(flat), 600), f32(flat), 1.0)
    line = rng.normal(scale=1.5e-3, size=(900, 3)); line[:, 0] = np.linspace(-0.1, 0.1, 900)        # a thin rod
    cases["rod"] = (moved(f32(line), 500, rot=0.01), f32(line), 1.0)
    dup = f32(rng.uniform(-0.05, 0.05, size=(600, 3)))
    dup = np.concatenate([dup, dup[::3], dup[::7]])                                                # exact duplicates
    cases["duplicates"] = (moved(dup, 800), dup, 1.0)
    g = np.linspace(-0.05, 0.05, 12)
    lat = f32(np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3))
    cases["lattice"] = (moved(lat, 1000, rot=0.01, shift=1e-3, noise=0.0), lat, 1.0)
    cases["tight_threshold"] = (moved(vol, 800), vol, 0.004)
    cases["few_targets"] = (f32(rng.uniform(-0.05, 0.05, size=(600, 3))), f32(rng.uniform(-0.05, 0.05, size=(3, 3))), 1.0)
    big = f32(rng.uniform(-0.1, 0.1, size=(12000, 3)))
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def run(names, th, extra=None, max_iter=60, hint=None):
        srcs = [cases[n][0] for n in names] + ([extra[0]] if extra else [])
        tgts = [cases[n][1] for n in names] + ([extra[1]] if extra else [])
        s_off = np.concatenate([[0], np.cumsum([s.shape[0] for s in srcs])]).astype(np.int32)
        t_off = np.concatenate([[0], np.cumsum([q.shape[0] for q in tgts])]).astype(np.int32)
        r = ci.icp_sweep(t(np.concatenate(srcs)), t(s_off), t(np.concatenate(tgts)), t(t_off),
                         t(np.arange(len(srcs), dtype=np.int32)), None, None, t(np.stack([np.eye(4)] * len(srcs))),
                         max_src_per_tile=int(np.diff(s_off).max()) if hint is None else hint, max_corr=th, max_iter=max_iter)
        torch.cuda.synchronize()
        return r, s_off, srcs, tgts

    # no size hint (max_src_per_tile = 0): the slice does not fit the shared memory sized for the default, so the
    # point state lives in the workspace spill area (unsorted points, no cache) -- same answers
    r, s_off, srcs, tgts = run(["sheet", "rod"], 1.0, max_iter=25, hint=0)
    for n, name in enumerate(["sheet", "rod"]):
        o = oracle.icp_p2p(srcs[n], tgts[n], 1.0, np.eye(4), max_iter=25, use_kdtree=True)
        assert int(r.iters[n]) == o["iters"] and np.array_equal(r.corr.cpu().numpy()[s_off[n]:s_off[n + 1]], o["corr"]), name
        assert np.abs(r.T[n].cpu().numpy() - o["T"]).max() <= 1e-5

    cases["far"] = (far[0], far[1], -1.0)
    r, s_off, srcs, tgts = run(["far", "sheet"], 1.0, max_iter=0)
    o = oracle.icp_p2p(far[0], far[1], 1.0, np.eye(4), max_iter=0, use_kdtree=True)
    assert np.array_equal(r.corr.cpu().numpy()[:s_off[1]], o["corr"]), "far: initial correspondences differ"
    r, s_off, srcs, tgts = run(["far"], 1.0, (moved(big, 9000), big), max_iter=0)
    assert np.array_equal(r.corr.cpu().numpy()[:s_off[1]], o["corr"]), "far (cluster variant): initial correspondences differ"
    for th in (1.0, 0.004):
        names = [n for n, c in cases.items() if c[2] == th]
        for extra in (None, (moved(big, 9000), big)):          # without / with a tile that switches on the cluster variant
            r, s_off, srcs, tgts = run(names, th, extra)
            corr, iters, T = r.corr.cpu().numpy(), r.iters.cpu().numpy(), r.T.cpu().numpy()
            for n in range(len(srcs)):
                o = oracle.icp_p2p(srcs[n], tgts[n], th, np.eye(4), max_iter=60, use_kdtree=True)
                name = names[n] if n < len(names) else "big"
                assert iters[n] == o["iters"], f"{name}: iterations {iters[n]} vs {o['iters']}"
                bad = np.nonzero(corr[s_off[n]:s_off[n + 1]] != o["corr"])[0]
                assert bad.size == 0, f"{name}: {bad.size} correspondences differ (cluster variant: {extra is not None})"
                assert np.abs(T[n] - o["T"]).max() <= 1e-5, name
