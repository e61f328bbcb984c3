#!/usr/bin/env python
"""Time the reference's OWN builds of f() on this host's cores (builds a/b): the unmodified sources
compiled in place by oracle/Makefile into oracle/_ref/shud_ref_serial (make shud) and shud_ref_omp (-D_OPENMP_ON -fopenmp,
the flags of the reference's Makefile:155-165), run on the three shipped basins with `--time REPS`.  Needs /root/reference
(the basins' text inputs): runs in the build container; bench.py --impl reference calls it when that tree is present.
The OpenMP build is reduced physics (no ET partition, no lakes - SURVEY.md 2.1): a timing baseline only.
    python tools/time_reference.py [reps] > profiles/r02_reference_cpu_timing.json"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
WORK = "/tmp/shud_ref_timing"
SIZES = {"ccw": 1147, "heihe": 1779, "qhh": 4773}


def run(reps=500):
    exe = {k: os.path.join(ROOT, "oracle", "_ref", f"shud_ref_{k}") for k in ("serial", "omp")}
    if not os.path.isdir(os.path.join(REF, "input")) or not all(os.path.exists(v) for v in exe.values()):
        return None
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(os.path.join(WORK, "input"))
    out = {"host_cores": os.cpu_count(), "reps": reps, "basins": {}}
    try:
        out["cpu"] = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        pass
    for b in SIZES:
        shutil.copytree(os.path.join(REF, "input", b), os.path.join(WORK, "input", b))
        para = os.path.join(WORK, "input", b, b + ".cfg.para")
        txt = open(para).read().splitlines()
        for i, ln in enumerate(txt):
            if b == "heihe" and ln.split() and ln.split()[0] == "END":
                txt[i] = "END\t9490"   # the shipped END exceeds the forcing record
        open(para, "w").write("\n".join(txt) + "\n")
        rec = {"Ne": SIZES[b]}
        for kind, env in (("serial", {}), ("omp", {"OMP_NUM_THREADS": str(os.cpu_count())})):
            r = subprocess.run([exe[kind], b, os.path.join(WORK, f"{b}.{kind}.bin"), "--time", str(reps)], cwd=WORK,
                               capture_output=True, text=True, errors="replace", env={**os.environ, **env})
            m = re.search(r"time_per_f_us=([0-9.eE+-]+) cells_per_s=([0-9.eE+-]+)", r.stdout)
            if r.returncode != 0 or not m:
                rec[kind] = {"error": (r.stdout + r.stderr)[-300:]}
                continue
            rec[kind] = {"us_per_f": float(m.group(1)), "cell_updates_per_s": float(m.group(2)),
                         "threads": int(env.get("OMP_NUM_THREADS", 1))}
        out["basins"][b] = rec
    return out


if __name__ == "__main__":
    res = run(int(sys.argv[1]) if len(sys.argv) > 1 else 500)
    print(json.dumps(res, indent=1) if res else json.dumps({"unavailable": "no /root/reference or oracle/_ref binaries here"}))
