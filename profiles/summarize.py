#!/usr/bin/env python
"""Turn ncu artefacts (brought back in gpurun_out/) into the small tracked summaries under profiles/.

    python profiles/summarize.py gpurun_out/prof_r01_sweep.ncu-rep gpurun_out/launches_r01.csv r01

writes profiles/<tag>_kernels.json (per-kernel raw metrics of one launch each), profiles/<tag>_launches.csv
(the --metrics gpu__time_duration.sum launch list) and prints the shares.  Needs the ncu CLI (here, no GPU)."""
import csv
import io
import json
import shutil
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum"]


def main(rep, launches, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        k = {"kernel": r[ix["Kernel Name"]].split("(")[0], "id": r[ix["ID"]]}
        for w in WANT:
            if w in ix:
                try:
                    k[w] = float(r[ix[w]].replace(",", ""))
                except ValueError:
                    k[w] = r[ix[w]]
                k[w + "__unit"] = units[ix[w]]
        out.append(k)
    json.dump(out, open(f"profiles/{tag}_kernels.json", "w"), indent=1)
    shutil.copy(launches, f"profiles/{tag}_launches.csv")
    # shares from the launch list
    tot = {}
    with open(launches) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> microseconds
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += v
    s = sum(v[1] for v in tot.values())
    print(f"{'kernel':60s} {'launches':>8s} {'avg us':>10s} {'share':>7s}")
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {n:8d} {v / n:10.1f} {100 * v / s:6.1f}%")
    for k in out:
        print(k["kernel"], {w: k.get(w) for w in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                  "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
                                                  "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")})


if __name__ == "__main__":
    main(*sys.argv[1:4])
