"""world_size-2 gloo tests (CPU): the frame partition and the all-gather of fitted poses of the
multi-GPU layer, with the CPU oracle standing in for the per-rank CUDA sweep (the oracle is
used here as the checker AND as the injected local compute, never by the product path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from autourdf_b200 import synth
        from autourdf_b200.dist import frame_partition, sharded_sweep
        from oracle import icp_oracle as O

        b = synth.make_config("wx200", n_frames=6)

        def run_local(sub):
            return O.masked_icp_sweep(sub.src, sub.src_off, sub.tgt, sub.tgt_off, sub.tile_frame, sub.box, sub.box_off,
                                      sub.init_T, nthreads=1)

        allr, local, (f0, f1) = sharded_sweep(b, run_local, device="cpu")
        full = run_local(b)
        ok = (np.array_equal(allr["T"], full["T"]) and np.array_equal(allr["iters"], full["iters"])
              and np.array_equal(allr["rmse"], full["rmse"]) and np.array_equal(allr["fitness"], full["fitness"]))
        parts = frame_partition(b, world)
        # real-pipeline mode: ONE frame transition, its K clusters split over the ranks, the
        # frame's cloud replicated; and the same split over the whole batch
        one = b.frame_slice(2, 3)
        allt, loct, (b0, b1) = sharded_sweep(one, run_local, device="cpu", by="tiles")
        full1 = run_local(one)
        ok = ok and np.array_equal(allt["T"], full1["T"]) and np.array_equal(allt["iters"], full1["iters"])
        ok = ok and loct["T"].shape[0] == b1 - b0 and 0 < b1 - b0 < one.n_tiles
        allb, _, _ = sharded_sweep(b, run_local, device="cpu", by="tiles")
        ok = ok and np.array_equal(allb["T"], full["T"]) and np.array_equal(allb["fitness"], full["fitness"])
        # frames dealt round-robin: gathered rows come back in the original tile order
        allr, locr, _ = sharded_sweep(b, run_local, device="cpu", by="frames_rr")
        ok = ok and np.array_equal(allr["T"], full["T"]) and np.array_equal(allr["iters"], full["iters"])
        ok = ok and np.array_equal(allr["rmse"], full["rmse"]) and 0 < locr["T"].shape[0] < b.n_tiles
        q.put((rank, ok, (f0, f1), parts, int(local["T"].shape[0])))
    except Exception as e:  # surface the failure instead of letting the parent wait for the queue
        q.put((rank, False, repr(e), None, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_sweep_equals_single_process_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    assert all(r[1] for r in res), f"gathered poses differ from the single-process sweep: {[r[2] for r in res]}"
    parts = res[0][3]
    assert parts[0][0] == 0 and parts[-1][1] == 5 and parts[0][1] == parts[1][0]     # contiguous cover of 5 frames
    assert [r[2] for r in res] == parts
    assert sum(r[4] for r in res) == 50                                                # 5 frames x 10 clusters


def test_tile_partition_and_tile_slice():
    """cluster-level sharding: contiguous balanced tile ranges; a tile slice carries exactly the frames its
    tiles refer to, re-indexed, and sweeping the slices reproduces the sweep of the whole batch"""
    sys.path.insert(0, ROOT)
    from autourdf_b200 import synth
    from autourdf_b200.dist import tile_partition
    from oracle import icp_oracle as O
    O.build()
    b = synth.make_config("wx200", n_frames=4)
    run = lambda s: O.masked_icp_sweep(s.src, s.src_off, s.tgt, s.tgt_off, s.tile_frame, s.box, s.box_off, s.init_T, nthreads=1)
    full = run(b)
    for world in (1, 2, 3, 7, 40):
        parts = tile_partition(b, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == b.n_tiles
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        pieces = []
        for a, c in parts:
            sub = b.tile_slice(a, c)
            assert sub.n_tiles == c - a and sub.tgt_off.shape[0] == sub.n_frames + 1
            if c > a:
                assert sub.tile_frame.min() == 0 and sub.tile_frame.max() == sub.n_frames - 1
                assert sub.n_frames == np.unique(b.tile_frame[a:c]).size
                pieces.append(run(sub))
        assert np.array_equal(np.concatenate([p["T"] for p in pieces]), full["T"])
        assert np.array_equal(np.concatenate([p["corr"] for p in pieces]), full["corr"])
    one = b.frame_slice(1, 2)
    sizes = [c - a for a, c in tile_partition(one, 4)]
    assert sum(sizes) == one.n_tiles and max(sizes) - min(sizes) <= 2


def test_frame_partition_properties():
    sys.path.insert(0, ROOT)
    from autourdf_b200 import synth
    from autourdf_b200.dist import frame_partition
    b = synth.make_config("wx200")
    for world in (1, 2, 3, 4, 8, 16):
        parts = frame_partition(b, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == b.n_frames
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert all(a <= c for a, c in parts)
        sizes = [c - a for a, c in parts]
        if world <= b.n_frames:
            assert max(sizes) - min(sizes) <= 2
    sub = b.frame_slice(*frame_partition(b, 2)[1])
    assert sub.tile_frame.min() == 0 and sub.tgt_off[0] == 0 and sub.src_off[0] == 0


def test_single_process_share_accepts_tensors_and_rejects_unsorted_frames():
    """world == 1: the local result may be torch tensors (the CUDA run_local returns device tensors);
    sharding by frames refuses a batch whose tiles are not frame-major instead of mis-slicing it"""
    sys.path.insert(0, ROOT)
    from autourdf_b200 import synth
    from autourdf_b200.dist import frame_partition, sharded_sweep
    from oracle import icp_oracle as O
    O.build()
    b = synth.make_config("wx200", n_frames=3)

    def run_local(sub):
        o = O.masked_icp_sweep(sub.src, sub.src_off, sub.tgt, sub.tgt_off, sub.tile_frame, sub.box, sub.box_off,
                               sub.init_T, nthreads=1)
        return {k: torch.from_numpy(np.asarray(v)) for k, v in o.items() if k in ("T", "fitness", "rmse", "iters")}

    allr, local, (f0, f1) = sharded_sweep(b, run_local)
    assert (f0, f1) == (0, b.n_frames) and isinstance(allr["T"], np.ndarray) and allr["T"].shape == (b.n_tiles, 4, 4)
    assert np.array_equal(allr["T"], local["T"].numpy())
    b.tile_frame = b.tile_frame[::-1].copy()
    with pytest.raises(ValueError):
        frame_partition(b, 2)


def test_frame_select_and_interleaved_partition():
    """round-robin frame sharding: every frame on exactly one rank, frame_select carries exactly those frames
    (re-indexed) and their tiles, and the sub-batch sweeps to the same poses as the whole batch"""
    from autourdf_b200 import synth
    from autourdf_b200.dist import frame_partition_interleaved
    from oracle import icp_oracle as O
    b = synth.make_config("wx200", n_frames=6)
    parts = frame_partition_interleaved(b, 3)
    assert sorted(np.concatenate(parts).tolist()) == list(range(b.n_frames))
    full = O.masked_icp_sweep(b.src, b.src_off, b.tgt, b.tgt_off, b.tile_frame, b.box, b.box_off, b.init_T)
    seen = np.zeros(b.n_tiles, dtype=bool)
    for fr in parts:
        sub, tiles = b.frame_select(fr)
        assert sub.n_frames == fr.size and sub.n_tiles == tiles.size and not seen[tiles].any()
        seen[tiles] = True
        assert np.array_equal(np.unique(sub.tile_frame), np.arange(fr.size))
        r = O.masked_icp_sweep(sub.src, sub.src_off, sub.tgt, sub.tgt_off, sub.tile_frame, sub.box, sub.box_off, sub.init_T)
        assert np.array_equal(r["T"], full["T"][tiles]) and np.array_equal(r["iters"], full["iters"][tiles])
    assert seen.all()
    empty, t = b.frame_select([])
    assert empty.n_tiles == 0 and t.size == 0
