#!/usr/bin/env python
"""Per-kernel sums of the launch list tools/nk_profile.py produces under ncu (gpu__time_duration.sum): GPU time per BDF
step by kernel, for the last `nsteps` steps (argument 2, default 6) that hold `nrhs` RHS calls (argument 3)."""
import csv, re, sys, collections
path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 18
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v = float(r[vi]) / (1e3 if r[ui] == "ns" else 1.0)
    n = re.sub(r"^void ", "", r[ki]).replace("<unnamed>::", "")
    n = re.sub(r"\(.*", "", n)
    seq.append((n, v))
idx = [i for i, (n, v) in enumerate(seq) if n.startswith("k_fused")]
tail = seq[idx[-nrhs] - 1:]
tot = collections.defaultdict(lambda: [0, 0.0])
for n, v in tail:
    tot[n][0] += 1
    tot[n][1] += v
S = sum(v for n, v in tail)
print(f"GPU time {S / nsteps:.1f} us per step, {len(tail) / nsteps:.1f} launches per step, {S / nrhs:.1f} us per RHS call")
for n, (c, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{n[:64]:64s} {c / nsteps:5.1f} /step {v / nsteps:8.1f} us/step {v / c:7.1f} us each")
