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
