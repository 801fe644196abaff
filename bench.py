#!/usr/bin/env python
"""bench.py -- cluster-ICP frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over the wx200_5 workload (C2: 5 sequences x 10 frames
= 45 frame transitions, 2048 points/frame, 20 clusters -> 900 (frame, cluster) tiles); one
"frame" is one complete masked_icp sweep of all clusters of one frame transition, run to the
reference's convergence rule.  Synthetic data (autourdf_b200.synth), float64 arithmetic.

  value     frames/s with inputs resident in HBM (CUDA events, max over ranks, L2 flushed
            between steps); at N GPUs every rank sweeps its own sequences (weak scaling)
            and the fitted poses are all-gathered over NCCL inside the timed region
  e2e       the same through the host-buffer C-ABI call (aurdf_icp_sweep_host): numpy in,
            numpy out, H2D + D2H copies inside the timed region (the call cuts the batch into
            three frame blocks on separate streams so copies and kernels overlap)
  roofline  the fused per-tile ICP kernel (icp_small_kernel: every wx200_5 tile is in its class):
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

# one JSON line on stdout: keep NCCL's version banner out of it
os.environ["NCCL_DEBUG"] = os.environ.get("AURDF_NCCL_DEBUG", "WARN")

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


def make_workload(rank=0):
    from autourdf_b200 import synth
    cfg = dict(synth.CONFIGS[WORKLOAD])
    return synth.make_batch(**cfg, seed=cfg["cid"] * 1000 + 17 * rank)


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
    from oracle import icp_oracle as O
    O.build()
    b = make_workload(0)
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
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": b.n_frames, "points_per_frame": b.meta["n_points"],
                   "clusters": b.n_clusters, "tiles_per_step": b.n_tiles},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": modes[mode]["cores"], "kind": "port",
                         "sample": REF_SAMPLE.format(w=WORKLOAD, f=b.n_frames, mode=mode),
                         "modes_frames_per_s": {k: m["value"] for k, m in modes.items()},
                         "host_threads": O.lib().orc_max_threads()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

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

    b = make_workload(rank)
    d = ci.batch_to_device(b, device=dev)
    max_src = int(np.diff(b.src_off).max())
    r0 = ci.icp_sweep(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"],
                      d["init_T"], max_src_per_tile=max_src)
    torch.cuda.synchronize()
    need = r0.needed_capacity()
    ntgt = r0.ntgt.cpu().numpy().astype(np.int64)
    iters = r0.iters.cpu().numpy().astype(np.int64)
    ns = np.diff(b.src_off).astype(np.int64)
    plan = ci.IcpSweep(b.n_tiles, b.src.shape[0], need + 64, max_src, device=dev)
    gathered = torch.empty((world,) + tuple(plan.out.T.shape), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    no_gather = os.environ.get("AURDF_BENCH_NO_GATHER") == "1"      # diagnostic only

    def step_eager():
        r = plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
        if world > 1 and not no_gather:   # the one exchange step of the path: fitted poses of every rank's sweep
            dist.all_gather_into_tensor(gathered, r.T)
        return r

    # N > 1: the step (5 kernels + the NCCL all-gather) is captured once in a CUDA graph, so a step costs
    # the host one launch instead of ~8 and the ranks do not drift apart on host jitter
    step, graph = step_eager, None
    if world > 1 and os.environ.get("AURDF_BENCH_GRAPH", "1") == "1":
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
            plan.run(d["src"], d["src_off"], d["tgt"], d["tgt_off"], d["tile_frame"], d["box"], d["box_off"], d["init_T"])
        torch.cuda.synchronize()
    L.aurdf_icp_profile_enable(0)
    import ctypes as C
    kms, kn = C.c_double(), C.c_int32()
    L.aurdf_icp_profile_collect(C.byref(kms), C.byref(kn))
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    frames_per_step = b.n_frames * world
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
    for _ in range(args.steps):
        host.run(*hin, out=out)
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d, d2h = host.copy_bytes()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_step * args.steps / float(te.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        # ---------------- roofline of the dominant kernel ----------------
        peak, peak_src = load_peaks()
        k_ms = kms.value / max(kn.value, 1)
        b_alg = float((28 * ns + 12 * ntgt + 128).sum())          # fused ICP tile, bytes per launch
        achieved = b_alg / (k_ms * 1e-3) / 1e9
        pairs = float((ns * ntgt * (iters + 1)).sum())             # distance evaluations per launch
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        # The roof that binds is instruction issue, not HBM: the float32 pre-filter of icp_small_kernel
        # issues 17 instructions per PAIR of targets (2 LDS.128, 6 packed f32x2, 2 LOP3, 4 VIMNMX,
        # 1 VIMNMX3, ~2 loop; SASS count), i.e. 8.5 per (source, target) evaluation; exact float64 work
        # is O(1) per point; peak = 148 SMs x 4 schedulers x 32 lanes per clock.
        instr_per_pair = 8.5
        issue_peak = 148 * 128 * sm_hz
        issue_rate = instr_per_pair * pairs / (k_ms * 1e-3)
        traffic, traffic_src = load_traffic()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "kernel": "icp_small_kernel", "kernel_ms": k_ms,
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
                   "sample": REF_SAMPLE.format(w=WORKLOAD, f=b.n_frames, mode=mode) + "; best of repeated runs",
                   "modes_frames_per_s": {k: m["value"] for k, m in modes.items()},
                   "host_threads": O.lib().orc_max_threads()}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": b.n_frames,
                       "points_per_frame": b.meta["n_points"], "clusters": b.n_clusters,
                       "tiles_per_step_per_gpu": b.n_tiles, "mean_icp_iters": float(iters.mean()),
                       "l2": "flushed between steps (256 MiB write)",
                       "parallelism": f"tiles sharded by sequence over {world} GPU(s), all-gather of poses",
                       "step_launch": "cuda_graph" if graph is not None else "eager"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(L.aurdf_icp_sweep_launches()) * args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
