#!/usr/bin/env python
"""bench.py -- cluster-ICP frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--scaling weak|strong] [--workload wx200_5|franka|allegro_hand|c5:<pts>x<clusters>]

A "step" is one pass of the hot path over the wx200_5 workload (C2: 5 sequences x 10 frames
= 45 frame transitions, 2048 points/frame, 20 clusters -> 900 (frame, cluster) tiles); one
"frame" is one complete masked_icp sweep of all clusters of one frame transition, run to the
reference's convergence rule.  Synthetic data (autourdf_b200.synth), float64 arithmetic.

  value     frames/s with inputs resident in HBM (CUDA events, max over ranks, L2 flushed
            between steps); at N GPUs every rank sweeps one wx200_5 batch (weak scaling: per-GPU
            work fixed) and the fitted poses are all-gathered over NCCL inside the timed region
  e2e       the same through the host-buffer C-ABI call (aurdf_icp_sweep_host): numpy in,
            numpy out, H2D + D2H copies inside the timed region (the call cuts the batch into
            three frame blocks on separate streams so copies and kernels overlap)
  roofline  the fused per-tile ICP kernel (icp_small2_kernel: every wx200_5 tile is in its class):
            algorithmic bytes / its measured duration vs the measured HBM peak
            (MEASURED_PEAKS.json), plus the instruction-issue fraction that actually binds it
  cpu_baseline / --impl reference: the CPU restatement (oracle/, k-d tree NN, OpenMP over
            tiles, all host threads) on the same workload -- the reference's own open3d path
            is not installable here
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# NCCL_DEBUG is the caller's to set (the driver reads NCCL's own log for the rank count); the JSON
# line is the LAST line this process prints on stdout.

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "wx200_5"
METRIC = "cluster_icp_frames_per_sec"
UNIT = "frames/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def load_traffic(kernel="icp_small_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the newest
    committed ncu --set full summary (profiles/*_kernels.json, written by profiles/summarize.py)."""
    import glob
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernels.json")), reverse=True):
        try:
            for k in json.load(open(path)):
                if kernel in k["kernel"]:
                    rd = k["dram__bytes_read.sum"] * mult[k["dram__bytes_read.sum__unit"]]
                    wr = k["dram__bytes_write.sum"] * mult[k["dram__bytes_write.sum__unit"]]
                    return rd + wr, os.path.basename(path)
        except Exception:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(rank=0, name=None):
    """the synthetic batch of a named BASELINE.json config, or of a C5 sweep point "c5:<points>x<clusters>[x<frames>]"
    (one sequence of <frames> transitions + 1 frames; default 10 transitions)"""
    from autourdf_b200 import synth
    name = name or WORKLOAD
    if name.startswith("c5:"):
        dims = [int(x) for x in name[3:].split("x")]
        n_pts, n_cl = dims[0], dims[1]
        n_tr = dims[2] if len(dims) > 2 else 10
        return synth.make_batch(n_points=n_pts, n_clusters=n_cl, n_seq=1, n_frames=n_tr + 1, dof=5, cid=5,
                                seed=5000 + 17 * rank)
    cfg = dict(synth.CONFIGS[name])
    return synth.make_batch(**cfg, seed=cfg["cid"] * 1000 + 17 * rank)


def config_dict(b, name, scaling, world):
    """identical in both arms (the driver compares them)"""
    per = "" if scaling == "strong" else "_per_gpu"
    return {"workload": name, "frames_per_step" + per: b.n_frames, "points_per_frame": b.meta["n_points"],
            "clusters": b.n_clusters, "tiles_per_step" + per: b.n_tiles,
            "l2": "flushed between steps (256 MiB write) on the GPU arm",
            "parallelism": (f"one {name} batch sharded by frames (round-robin) over {world} GPU(s)" if scaling == "strong" else
                            f"every GPU sweeps one {name} batch ({world} GPU(s))") + ", one all-gather of poses"}


def finish(world, line=None):
    """bounded exit: tear the process group down under a watchdog, THEN print the JSON line (rank 0), so it is
    the last line on stdout even when NCCL_DEBUG=INFO logs the communicator teardown, and leave"""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(20.0)
        if line is not None:
            time.sleep(0.5)          # let the other ranks' teardown lines through first
    if line is not None:
        print(line, flush=True)
    sys.stderr.flush()
    os._exit(0)


def cpu_sweep(O, b, nthreads=0):
    return O.masked_icp_sweep(b.src, b.src_off, b.tgt, b.tgt_off, b.tile_frame, b.box, b.box_off, b.init_T,
                              use_kdtree=True, nthreads=nthreads)


def _time(fn, min_reps, budget_s, max_reps=200):
    fn()
    best, reps, t_all = 1e30, 0, time.perf_counter()
    while reps < min_reps or (time.perf_counter() - t_all < budget_s and reps < max_reps):
        t1 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t1)
        reps += 1
    return best, reps


def cpu_modes(O, b, budget_s=4.0):
    """frames/s of the CPU restatement under three threading structures:
      reference_serial    tiles one after the other on one thread (the Python loop at
                          cluster_icp.py:131 around a single-threaded open3d)
      reference_openmp    tiles one after the other, OpenMP over the source points inside the
                          correspondence search on every host thread (open3d's own structure)
      tile_parallel       OpenMP over tiles on every host thread -- NOT the reference's structure,
                          the strongest CPU port we could write; reported for context."""
    cores = O.use_all_host_threads()
    out = {}
    O.set_reference_threading(False)
    t, r = _time(lambda: cpu_sweep(O, b, 1), 2, budget_s)
    out["reference_serial"] = dict(value=b.n_frames / t, cores=1, reps=r)
    O.set_reference_threading(True)
    t, r = _time(lambda: cpu_sweep(O, b, 0), 1, budget_s, max_reps=20)
    out["reference_openmp"] = dict(value=b.n_frames / t, cores=cores, reps=r)
    O.set_reference_threading(False)
    t, r = _time(lambda: cpu_sweep(O, b, 0), 3, budget_s)
    out["tile_parallel"] = dict(value=b.n_frames / t, cores=cores, reps=r)
    return out


REF_SAMPLE = ("full {w} batch ({f} frame transitions) per step; restated open3d point-to-point ICP "
              "(oracle/icp_oracle.c: k-d tree NN, float64), tiles processed one after the other as the reference's "
              "Python loop does, threading = the faster of single-thread and open3d-style OpenMP over source points "
              "({mode}); an OpenMP-over-tiles port is listed beside it for context")


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path in the reference's own
    structure (see cpu_modes).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from oracle import icp_oracle as O
    O.build()
    b = make_workload(0, args.workload)
    modes = cpu_modes(O, b, budget_s=2.0)
    mode = "reference_serial" if modes["reference_serial"]["value"] >= modes["reference_openmp"]["value"] else "reference_openmp"
    inner = mode == "reference_openmp"
    O.set_reference_threading(inner)
    nthr = 0 if inner else 1
    for _ in range(max(args.warmup, 1)):
        cpu_sweep(O, b, nthr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sweep(O, b, nthr)
    dt = time.perf_counter() - t0
    O.set_reference_threading(False)
    v = b.n_frames * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(b, args.workload, args.scaling, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": modes[mode]["cores"], "kind": "port",
                         "sample": REF_SAMPLE.format(w=args.workload, f=b.n_frames, mode=mode),
                         "modes_frames_per_s": {k: m["value"] for k, m in modes.items()},
                         "all_core_port_frames_per_s": modes["tile_parallel"]["value"],
                         "host_threads": O.lib().orc_max_threads()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's contract): every GPU sweeps one batch of --workload; strong: ONE "
                         "batch of --workload sharded over the GPUs by frame blocks (autourdf_b200.dist.ShardedSweep)")
    ap.add_argument("--workload", default=WORKLOAD)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C

    import torch
    import torch.distributed as dist
    from autourdf_b200 import _lib
    from autourdf_b200 import cluster_icp as ci

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    strong = args.scaling == "strong"

    # weak: every rank sweeps the SAME batch (per-GPU work exactly fixed as N grows, the definition of weak scaling;
    # ICP work is data-dependent, so per-rank seeds would give every rank a different amount of work --
    # AURDF_BENCH_RANK_SEEDS=1 selects that variant, profiles/r02_scaling.md has both); strong: the same batch on
    # every rank, of which it keeps its share
    rank_seeds = os.environ.get("AURDF_BENCH_RANK_SEEDS") == "1"
    b_all = make_workload(rank if (rank_seeds and not strong) else 0, args.workload)
    sharded = None
    if strong:
        from autourdf_b200.dist import ShardedSweep
        sharded = ShardedSweep(b_all, dev, by=os.environ.get("AURDF_BENCH_SHARD", "frames_rr"))
        b = sharded.sub
    else:
        b = b_all
    have = b.n_tiles > 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    ns = np.diff(b.src_off).astype(np.int64)
    if strong:
        plan, d = sharded.plan, sharded.d
        r0 = plan.out if have else None
        gathered = None
    else:
        d = ci.batch_to_device(b, device=dev)
        max_src = int(ns.max())
        r0 = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"],
                          d["init_T"], max_src_per_tile=max_src)
        torch.cuda.synchronize()
        plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], r0.needed_capacity() + 64, max_src, device=dev)
        gathered = torch.empty((world,) + tuple(plan.out.T.shape), dtype=torch.float64, device=dev) if world > 1 else None
    if strong:
        sharded.run()
    torch.cuda.synchronize()
    ntgt = r0.ntgt.cpu().numpy().astype(np.int64) if have else np.zeros(0, np.int64)
    iters = r0.iters.cpu().numpy().astype(np.int64) if have else np.zeros(0, np.int64)

    check = None
    if strong:
        # the gathered poses of the sharded sweep must be bit-identical to ONE GPU's sweep of the
        # whole batch (rank 0 runs it once, outside the timed region)
        if rank == 0:
            dall = ci.batch_to_device(b_all, device=dev)
            rf = ci.icp_sweep(dall["src"], dall["src_off"], dall["tgt"], dall["tgt_off"], dall["tile_frame"], dall["box"],
                              dall["box_off"], dall["init_T"], max_src_per_tile=int(np.diff(b_all.src_off).max()))
            torch.cuda.synchronize()
            check = bool(np.array_equal(sharded.poses(), rf.T.cpu().numpy()))
            assert check, "sharded sweep differs from the single-GPU sweep"
            del dall, rf

    no_gather = os.environ.get("AURDF_BENCH_NO_GATHER") == "1"      # diagnostic only

    def step_eager():
        if strong:
            return sharded.run()
        r = plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
        if world > 1 and not no_gather:   # the one exchange step of the path: fitted poses of every rank's sweep
            dist.all_gather_into_tensor(gathered, r.T)
        return r

    # Optional (AURDF_BENCH_GRAPH=1): the step (kernels + the NCCL all-gather) captured once in a CUDA graph.
    # Off by default: it bought 2 % at 4 GPUs and a live graph holding NCCL work must be destroyed before the
    # process group is.
    step, graph = step_eager, None
    if world > 1 and os.environ.get("AURDF_BENCH_GRAPH", "0") == "1":
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step_eager()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step_eager()
            step = graph.replay
        except Exception as e:                      # capture not possible here: run eagerly
            print(f"[bench] CUDA graph capture failed on rank {rank}: {e!r}; running eagerly", file=sys.stderr)
            step, graph = step_eager, None
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---------------- device-resident timing ----------------
    for i in range(args.warmup):
        flush.fill_(i & 0xFF)
        step()
    barrier()
    if graph is None:
        L.aurdf_icp_profile_enable(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)        # L2 flush, outside the timed bracket
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    dev_ms = sum(a.elapsed_time(b_) for a, b_ in ev)
    if graph is not None:            # the library's per-launch events cannot live inside a graph: time the
        L.aurdf_icp_profile_enable(1)  # dominant kernel on a few eager launches after the timed region
        for i in range(min(args.steps, 20)):
            flush.fill_(i & 0xFF)
            if have:
                plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
        torch.cuda.synchronize()
        del graph                    # before the process group goes away
        step = step_eager
        torch.cuda.synchronize()
        graph_used = True
    else:
        graph_used = False
    L.aurdf_icp_profile_enable(0)
    kms, kn = C.c_double(), C.c_int32()
    L.aurdf_icp_profile_collect(C.byref(kms), C.byref(kn))
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    frames_per_step = b_all.n_frames if strong else b.n_frames * world
    value = frames_per_step * args.steps / (dev_ms_max * 1e-3)

    # ---------------- end to end through the host-buffer C ABI ----------------
    # numpy arrays backed by page-locked memory (the contract's "pinned host memory"): the library
    # copies them without staging; pageable arrays work too, through its own pinned staging.
    host = ci.HostSweep(local_rank)
    keep = []

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(t)
        return t.numpy()

    h2d = d2h = 0
    e2e_s = 0.0
    if have:
        hin = [pinned(x) for x in (b.src, b.src_off, b.tgt, b.tgt_off, b.tile_frame, b.box, b.box_off, b.init_T)]
        B, N = b.n_tiles, b.src.shape[0]
        out = dict(T=pinned(np.empty((B, 4, 4))), world=pinned(np.empty((N, 3))), corr=pinned(np.empty(N, np.int32)),
                   fitness=pinned(np.empty(B)), rmse=pinned(np.empty(B)), iters=pinned(np.empty(B, np.int32)),
                   ntgt=pinned(np.empty(B, np.int32)))
        for _ in range(args.warmup):
            host.run(*hin, out=out)
        assert np.array_equal(out["iters"], iters.astype(np.int32)), "host path disagrees with the device path"
    barrier()
    t0 = time.perf_counter()
    if have:
        for _ in range(args.steps):
            host.run(*hin, out=out)
    e2e_s = time.perf_counter() - t0
    barrier()
    if have:
        h2d, d2h = host.copy_bytes()
    te = torch.tensor([e2e_s, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        tsum = te.clone()
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if strong:                   # strong: bytes of the whole job; weak: per GPU (as the config is)
            h2d, d2h = int(tsum[1].item()), int(tsum[2].item())
    e2e_value = frames_per_step * args.steps / float(te[0].item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        # ---------------- roofline of the dominant kernel ----------------
        peak, peak_src = load_peaks()
        k_ms = kms.value / max(kn.value, 1)
        b_alg = float((28 * ns + 12 * ntgt + 128).sum())          # fused ICP tile, bytes per launch
        achieved = b_alg / (k_ms * 1e-3) / 1e9
        pairs = float((ns * ntgt * (iters + 1)).sum())             # distance evaluations per launch
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        # The roof that binds is instruction issue, not HBM: the float32 pre-filter of icp_small2_kernel
        # issues 17 instructions per PAIR of targets (2 LDS.128, 6 packed f32x2, 2 LOP3, 4 VIMNMX,
        # 1 VIMNMX3, ~2 loop; SASS count), i.e. 8.5 per (source, target) evaluation; exact float64 work
        # is O(1) per point; peak = 148 SMs x 4 schedulers x 32 lanes per clock.
        instr_per_pair = 8.5
        issue_peak = 148 * 128 * sm_hz
        issue_rate = instr_per_pair * pairs / (k_ms * 1e-3)
        small = bool(ns.size and ns.max() <= 384 and ntgt.max() <= 760)
        kname = "icp_small2_kernel" if small else "icp_grid_kernel"
        traffic, traffic_src = load_traffic(kname)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "kernel": kname, "kernel_ms": k_ms,
                    "kernel_share_of_step": k_ms / (dev_ms / args.steps),
                    "algorithmic_bytes_per_launch": b_alg,
                    "binding_roof": {"bound": "instruction_issue_and_iteration_latency",
                                     "pair_evals_per_launch": pairs, "instr_per_pair": instr_per_pair,
                                     "achieved_lane_instr_per_s": issue_rate, "peak_lane_instr_per_s": issue_peak,
                                     "frac": issue_rate / issue_peak,
                                     "note": "launch length is set by the slowest tile (max ICP iterations vs mean): "
                                             f"{int(iters.max())} vs {float(iters.mean()):.1f}"}}
        # ---------------- CPU baseline on this box's host cores (N = 1 only) ----------------
        cpu = None
        if world == 1 and os.environ.get("AURDF_BENCH_SKIP_CPU") != "1":   # (the skip is for quick A/B runs only)
            from oracle import icp_oracle as O
            O.build()
            modes = cpu_modes(O, b, budget_s=5.0)
            mode = "reference_serial" if modes["reference_serial"]["value"] >= modes["reference_openmp"]["value"] else "reference_openmp"
            cpu = {"value": modes[mode]["value"], "unit": UNIT, "cores": modes[mode]["cores"], "kind": "port",
                   "sample": REF_SAMPLE.format(w=args.workload, f=b.n_frames, mode=mode) + "; best of repeated runs",
                   "modes_frames_per_s": {k: m["value"] for k, m in modes.items()},
                   "all_core_port_frames_per_s": modes["tile_parallel"]["value"],
                   "host_threads": O.lib().orc_max_threads()}
        line = json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(b_all, args.workload, args.scaling, world),
            "detail": {"mean_icp_iters": float(iters.mean()) if iters.size else 0.0,
                       "max_icp_iters": int(iters.max()) if iters.size else 0,
                       "step_launch": "cuda_graph" if graph_used else "eager",
                       "sharded_equals_single_gpu": check,
                       "rank0_tiles": int(b.n_tiles)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(L.aurdf_icp_sweep_launches()) * args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        })
    host.close()
    finish(world, line if rank == 0 else None)


if __name__ == "__main__":
    main()
