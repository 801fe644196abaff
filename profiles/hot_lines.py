#!/usr/bin/env python
"""Aggregate the warp-stall samples and executed instructions of the kernel in an ncu report by SOURCE LINE
(needs -lineinfo at compile time and --import-source on at capture time; runs here, no GPU).
    python profiles/hot_lines.py <report.ncu-rep> [top] [kernel-name regex]"""
import collections, csv, io, subprocess, sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
sel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr, data = rows[h], rows[h + 1:]
iline, isrc = hdr.index("Line No"), hdr.index("Source")
isamp, iex, ith = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), ""])
cur_file = ""
for r in rows[:h]:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
seen_fn = 0
for r in data:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        seen_fn += 1
        continue
    if len(r) <= ith or r[iline] == "Line No":
        continue
    a = agg[(cur_file + ":" + r[iline]) if r[iline].strip() else ""]
    a[0] += num(r[isamp]); a[1] += num(r[iex]); a[2] += num(r[ith]); a[4] = r[isrc]
    for i in stall:
        if r[i]:
            a[3][hdr[i][6:]] += num(r[i])
ts, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{ts} samples, {te / 1e6:.1f} M warp instructions")
print("| line | samples % | warp instr % | active lanes | top stalls | source |\n|---|---|---|---|---|---|")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in a[3].most_common(3))
    print(f"| {line} | {100 * a[0] / ts:.1f} | {100 * a[1] / te:.1f} | {a[2] / max(a[1], 1):.1f} | {st} | `{a[4].strip()[:90]}` |")
