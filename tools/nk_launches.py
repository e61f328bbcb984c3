#!/usr/bin/env python
"""Per-kernel sums of the launch list tools/nk_profile.py produces under ncu: GPU time per BDF step by kernel, for the
last `nsteps` steps (argument 2, default 6) that hold `nrhs` RHS calls (argument 3).  With
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
the DRAM traffic and bandwidth of each kernel are listed as well."""
import csv, re, sys, collections
path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 18
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ii, ki, mi, vi, ui = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
launch = collections.OrderedDict()
for r in rows[1:]:
    n = re.sub(r"^void ", "", r[ki]).replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n)
    e = launch.setdefault(r[ii], {"name": n, "us": 0.0, "bytes": 0.0})
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        e["us"] = v / (1e3 if r[ui] == "ns" else 1.0) if r[ui] in ("ns", "us") else v * 1e3
    elif r[mi].startswith("dram__bytes"):
        e["bytes"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
seq = list(launch.values())
idx = [i for i, e in enumerate(seq) if e["name"].startswith("k_fused")]
tail = seq[idx[-nrhs] - 1:]
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in tail:
    t = tot[e["name"]]
    t[0] += 1; t[1] += e["us"]; t[2] += e["bytes"]
S = sum(e["us"] for e in tail)
print(f"GPU time {S / nsteps:.1f} us per step, {len(tail) / nsteps:.1f} launches per step, {S / nrhs:.1f} us per RHS call")
for n, (c, v, b) in sorted(tot.items(), key=lambda x: -x[1][1]):
    extra = f" {b / c / 1e6:7.1f} MB  {b / v / 1e3:6.0f} GB/s" if b > 0 else ""
    print(f"{n[:56]:56s} {c / nsteps:5.1f} /step {v / nsteps:8.1f} us/step {v / c:7.1f} us each{extra}")
