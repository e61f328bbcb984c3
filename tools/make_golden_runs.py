#!/usr/bin/env python
"""Generate tests/golden/<basin>.run.npz: the land-surface inputs of a 30-day window of each shipped basin, as the
UNMODIFIED reference replays them (oracle/ref_driver.cpp --land-seq: updateAllTimeSeries + updateforcing + ET once per
SolverStep, exactly the cadence of the reference's time loop, src/Model/shud.cpp:91-109 with ETStep >= SolverStep).
Only the per-step INPUTS of the per-cell land step are kept (station rows, LAI / melt-factor class values, solar
samples) plus the reference's outputs of the LAST step (pin of the whole bucket sequence): a few hundred kB per basin.
The window starts three days before the first substantial rain after START, so the run contains a flow event.
Runs in the build container only (needs /root/reference and oracle/_ref/shud_ref_serial)."""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shud_up_b200 import snapshot  # noqa: E402
from tools.make_golden import EXE, OUT, REF, rainy_minute  # noqa: E402

WORK = "/tmp/shud_golden_runs"
DAYS = 30


def main():
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(os.path.join(WORK, "input"))
    for b in ("ccw", "heihe", "qhh"):
        shutil.copytree(os.path.join(REF, "input", b), os.path.join(WORK, "input", b))
        para = os.path.join(WORK, "input", b, b + ".cfg.para")
        txt = open(para).read().splitlines()
        cfg = {}
        for i, ln in enumerate(txt):
            w = ln.split()
            if b == "heihe" and w and w[0] == "END":
                txt[i] = "END\t9490"     # the shipped END exceeds the forcing record
            if len(w) >= 2:
                cfg[w[0]] = w[1]
        open(para, "w").write("\n".join(txt) + "\n")
        start = float(cfg["START"]) * 1440.0
        dt = float(cfg["MAX_SOLVER_STEP"])
        t0 = max(start, np.floor(rainy_minute(b, start) / 1440.0) * 1440.0 - 3 * 1440.0)
        n = int(round(DAYS * 1440.0 / dt))
        binf = os.path.join(WORK, f"{b}.run.bin")
        cmd = [EXE, b, binf, "--land-seq", str(n), "--land-t0", str(t0), "--land-dt", str(dt), "--land-stride", str(10 ** 9)]
        r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
        if r.returncode != 0:
            sys.exit(f"reference run failed: {cmd}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        snap = snapshot.read_bin(binf)
        keep = {k: v for k, v in snap.items() if k.startswith("land_") or k.startswith("lseq_")}
        keep["run_cfg"] = np.array([float(cfg["RELTOL"]), float(cfg["ABSTOL"]), float(cfg["INIT_SOLVER_STEP"]), dt, t0, n])
        keep["_cmd"] = np.array(" ".join(cmd[1:]))
        out = os.path.join(OUT, f"{b}.run.npz")
        np.savez_compressed(out, **keep)
        print(b, "t0", t0, "dt", dt, "steps", n, "->", out, os.path.getsize(out) // 1024, "kB")


if __name__ == "__main__":
    main()
