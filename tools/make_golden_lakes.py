#!/usr/bin/env python
"""Golden fixture for the river / lake couplings no shipped basin exercises (parity gaps): reaches that
flow INTO a lake (down <= -4, src/ModelData/MD_Lake.cpp:46-54 -> QLakeRivIn, MD_RiverFlux.cpp:16-24), more than one
lake, and the outlet codes -1 / -2 (MD_RiverFlux.cpp:36-48).  The qhh inputs are copied to a scratch directory and
patched as TEXT - the lake of qhh is split into two lakes (second half of its cells, a second bathymetry table), and
the 45 outlet reaches get down = -4 (into lake 1), -5 (into lake 2), -1, -2 or keep -3 in turn - then the UNMODIFIED
reference (oracle/_ref/shud_ref_serial) runs f() twice on a randomised state, as tools/make_golden.py does.
Writes tests/golden/qhh.lakes6.npz: every array that differs from tests/golden/qhh.mesh.npz plus the dynamic ones.

Runs only in the build container (needs /root/reference); the fixture it writes is what travels."""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shud_up_b200 import snapshot  # noqa: E402
from tools.make_golden import DYNAMIC_ELE, EXE, OUT, REF, STATIC_PREFIX, STATIC_SCALARS, rainy_minute  # noqa: E402

WORK = "/tmp/shud_golden_lakes"


def patch_inputs(d):
    att = open(os.path.join(d, "qhh.sp.att")).read().splitlines()
    rows = [ln.split() for ln in att[2:]]
    lake_rows = [k for k, r in enumerate(rows) if int(r[8]) > 0]
    for k in lake_rows[len(lake_rows) // 2:]:
        rows[k][8] = "2"
    open(os.path.join(d, "qhh.sp.att"), "w").write("\n".join(att[:2] + ["\t".join(r) for r in rows]) + "\n")
    bathy = open(os.path.join(d, "qhh.lake.bathy")).read().rstrip("\n")
    bathy += "\n3\t3\nINDEX\tDEPTH AREA\n1   3152 1900000000\n2\t3161 2100000000\n3\t3228 2300000000\n"
    open(os.path.join(d, "qhh.lake.bathy"), "w").write(bathy)
    riv = open(os.path.join(d, "qhh.sp.riv")).read().splitlines()
    n = int(riv[0].split()[0])
    codes, k = ["-4", "-5", "-1", "-2", "-3"], 0
    for i in range(2, 2 + n):
        f = riv[i].split()
        if int(f[1]) < 0:
            f[1] = codes[k % len(codes)]
            k += 1
            riv[i] = "\t".join(f)
    open(os.path.join(d, "qhh.sp.riv"), "w").write("\n".join(riv) + "\n")
    return k


def main():
    if not os.path.exists(EXE):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(os.path.join(WORK, "input"))
    d = os.path.join(WORK, "input", "qhh")
    shutil.copytree(os.path.join(REF, "input", "qhh"), d)
    nout = patch_inputs(d)
    start = None
    for ln in open(os.path.join(d, "qhh.cfg.para")):
        if ln.split() and ln.split()[0] == "START":
            start = float(ln.split()[1]) * 1440.0
    binf = os.path.join(WORK, "qhh.lakes6.bin")
    cmd = [EXE, "qhh", binf, "--state", "rand:6", "--t", str(rainy_minute("qhh", start)), "--mutate", "frozen,ss"]
    r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
    tail = [ln for ln in r.stdout.splitlines() if "[shud_ref]" in ln]
    if r.returncode != 0 or not tail:
        sys.exit(f"reference run failed: {cmd}\n{r.stdout[-3000:]}\n{r.stderr[-2000:]}")
    print(" ".join(cmd[1:]), "->", tail[-1])
    snap = snapshot.read_bin(binf)
    base = dict(np.load(os.path.join(OUT, "qhh.mesh.npz")))
    is_static = lambda k: (k in STATIC_SCALARS or k.startswith(STATIC_PREFIX)) and k not in DYNAMIC_ELE
    dyn = {}
    for k, v in snap.items():
        if is_static(k) and k in base and np.array_equal(base[k], v):
            continue
        dyn[k] = v
    dyn["_cmd"] = np.array("patched qhh inputs (tools/make_golden_lakes.py): " + " ".join(cmd[1:]))
    np.savez_compressed(os.path.join(OUT, "qhh.lakes6.npz"), **dyn)
    print("Nl", snap["Nl"], "outlets patched", nout, "toLake>=0:", int((snap["riv_toLake"] >= 0).sum()),
          "down codes:", {int(c): int((snap["riv_down"] == c).sum()) for c in (-1, -2, -3, -4, -5)},
          "QLakeRivIn", snap.get("QLakeRivIn"))
    print("static overrides:", sorted(k for k in dyn if is_static(k)))


if __name__ == "__main__":
    main()
