#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference RHS.

Runs in the build container only (needs /root/reference and oracle/_ref/shud_ref_serial,
built by `make -C oracle ref`).  For each case it runs the reference driver
(oracle/ref_driver.cpp) from a scratch copy of the basin inputs and stores the snapshot:
  tests/golden/<basin>.mesh.npz        static SoA arrays of the basin (shared by its cases)
  tests/golden/<basin>.<case>.npz      forcing, carried state, y, reference ydot + flux arrays,
                                       and any static array a mutation changed
heihe is run with END patched to 9490: the shipped END 9861 exceeds its 9496-day forcing and
the reference aborts.
"""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from shud_up_b200 import snapshot  # noqa: E402

REF = "/root/reference"
EXE = os.path.join(ROOT, "oracle", "_ref", "shud_ref_serial")
WORK = "/tmp/shud_golden_work"
OUT = os.path.join(ROOT, "tests", "golden")

STATIC_PREFIX = ("ele_", "riv_", "seg_", "lake_")
STATIC_SCALARS = ("Ne", "Nr", "Ns", "Nl", "close_boundary", "lakeon")
DYNAMIC_ELE = ("ele_yBC", "ele_QBC", "ele_u_satn", "riv_yBC", "riv_qBC")


def rainy_minute(basin, start_min):
    a = np.loadtxt(os.path.join(REF, "input", basin, "forcing.csv"), skiprows=2)
    ok = np.where((a[:, 1] > 0.5) & (a[:, 2] > 3.0) & (a[:, 5] > 100.0) & (a[:, 0] * 1440 > start_min + 1440))[0]
    return float(round(a[ok[0], 0] * 1440.0)) + 5.0  # inside the rainy record, not on its edge


CASES = [
    # basin, case name, state, time ("start" | "rain"), mutations
    ("ccw", "ic", "ic", "start", ""),
    ("ccw", "rand1", "rand:1", "rain", ""),
    ("ccw", "mut2", "rand:2", "rain", "openbnd,frozen,ss,ebc,rbc,down4"),
    ("heihe", "ic", "ic", "start", ""),
    ("heihe", "rand3", "rand:3", "rain", ""),
    ("qhh", "ic", "ic", "start", ""),
    ("qhh", "rand4", "rand:4", "rain", ""),
    ("qhh", "mut5", "rand:5", "rain", "openbnd,frozen,ss,ebc,rbc"),
]


def main():
    if not os.path.exists(EXE):
        sys.exit("build oracle/_ref first: make -C oracle ref")
    shutil.rmtree(WORK, ignore_errors=True)
    os.makedirs(os.path.join(WORK, "input"))
    os.makedirs(OUT, exist_ok=True)
    starts = {}
    for b in ("ccw", "heihe", "qhh"):
        shutil.copytree(os.path.join(REF, "input", b), os.path.join(WORK, "input", b))
        para = os.path.join(WORK, "input", b, b + ".cfg.para")
        txt = open(para).read().splitlines()
        for i, ln in enumerate(txt):
            if b == "heihe" and ln.split() and ln.split()[0] == "END":
                txt[i] = "END\t9490"
            if ln.split() and ln.split()[0] == "START":
                starts[b] = float(ln.split()[1]) * 1440.0
        open(para, "w").write("\n".join(txt) + "\n")
    bases = {}
    for basin, case, state, when, mut in CASES:
        binf = os.path.join(WORK, f"{basin}.{case}.bin")
        cmd = [EXE, basin, binf, "--state", state]
        if when == "rain":
            cmd += ["--t", str(rainy_minute(basin, starts[basin]))]
        if mut:
            cmd += ["--mutate", mut]
        r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
        tail = [ln for ln in r.stdout.splitlines() if "[shud_ref]" in ln]
        if r.returncode != 0 or not tail:
            sys.exit(f"reference run failed: {cmd}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        print(" ".join(cmd[1:]), "->", tail[-1])
        snap = snapshot.read_bin(binf)
        is_static = lambda k: (k in STATIC_SCALARS or k.startswith(STATIC_PREFIX)) and k not in DYNAMIC_ELE
        if basin not in bases:
            bases[basin] = {k: v for k, v in snap.items() if is_static(k)}
            np.savez_compressed(os.path.join(OUT, f"{basin}.mesh.npz"), **bases[basin])
        base = bases[basin]
        dyn = {}
        for k, v in snap.items():
            if is_static(k) and k in base and np.array_equal(base[k], v):
                continue
            dyn[k] = v
        dyn["_cmd"] = np.array(" ".join(cmd[1:]))
        np.savez_compressed(os.path.join(OUT, f"{basin}.{case}.npz"), **dyn)
    # forcing sequence for a short full run (tests/test_integrator_*.py): 48 hourly land-surface steps of ccw
    binf = os.path.join(WORK, "ccw.fseq.bin")
    cmd = [EXE, "ccw", binf, "--forcing-seq", "48"]
    r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
    if r.returncode != 0:
        sys.exit(f"reference run failed: {cmd}\n{r.stdout[-2000:]}")
    snap = snapshot.read_bin(binf)
    keep = {k: v for k, v in snap.items() if k.startswith("fseq_") and k not in ("fseq_fu_Surf", "fseq_fu_Sub")}
    keep["_cmd"] = np.array(" ".join(cmd[1:]))
    np.savez_compressed(os.path.join(OUT, "ccw.fseq.npz"), **keep)
    # land-surface step (updateforcing + ET) sequences: per-step inputs and outputs of the reference itself
    # (ccw: 30 hourly steps through a rain / snow / melt event with terrain radiation; qhh: 8 steps, lake cells,
    # 3-hourly forcing so that three steps share one forcing interval)
    for basin, n, t0 in (("ccw", 30, 4254000), ("qhh", 8, 225120)):
        binf = os.path.join(WORK, f"{basin}.land.bin")
        cmd = [EXE, basin, binf, "--land-seq", str(n), "--land-t0", str(t0)]
        r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
        if r.returncode != 0:
            sys.exit(f"reference run failed: {cmd}\n{r.stdout[-2000:]}")
        snap = snapshot.read_bin(binf)
        keep = {k: v for k, v in snap.items() if k.startswith("land_") or k.startswith("lseq_")}
        keep["_cmd"] = np.array(" ".join(cmd[1:]))
        np.savez_compressed(os.path.join(OUT, f"{basin}.land.npz"), **keep)
    # frozen-soil factors (CRYOSPHERE = 1, switched on in memory with a -6 K temperature offset: no shipped basin
    # uses them): 800 hourly steps = 33 days, so that both running-mean windows (7 and 28 days) wrap; outputs kept
    # every 40th step, and only the ones the switch affects
    binf = os.path.join(WORK, "ccw.cryo.bin")
    cmd = [EXE, "ccw", binf, "--land-seq", "800", "--land-t0", "4254000", "--land-stride", "40", "--mutate", "cryo"]
    r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
    if r.returncode != 0:
        sys.exit(f"reference run failed: {cmd}\n{r.stdout[-2000:]}")
    snap = snapshot.read_bin(binf)
    kept_out = {"lseq_fu_Surf", "lseq_fu_Sub", "lseq_t_temp", "lseq_yEleSnow", "lseq_qEleNetPrep"}
    all_out = {"lseq_" + n for n in ("qElePrep", "qPotEvap", "qPotTran", "qEleETP", "t_lai", "t_temp", "t_mf", "qEleNetPrep",
                                     "qEleE_IC", "yEleSnow", "yEleIS", "fu_Surf", "fu_Sub", "rn_factor")}
    keep = {k: v for k, v in snap.items()
            if (k.startswith("land_") or k.startswith("lseq_")) and (k not in all_out or k in kept_out)}
    keep["_cmd"] = np.array(" ".join(cmd[1:]))
    np.savez_compressed(os.path.join(OUT, "ccw.cryo.npz"), **keep)
    # the reference's own checkpoint writer (Model_Data::PrintInit) on a random qhh state: pin of shud_b200_format_ic
    binf, txtf = os.path.join(WORK, "qhh.ic.bin"), os.path.join(WORK, "qhh.ic.txt")
    cmd = [EXE, "qhh", binf, "--state", "rand:4", "--print-init", txtf]
    r = subprocess.run(cmd, cwd=WORK, capture_output=True, text=True, errors="replace")
    if r.returncode != 0:
        sys.exit(f"reference run failed: {cmd}\n{r.stdout[-2000:]}")
    snap = snapshot.read_bin(binf)
    np.savez_compressed(os.path.join(OUT, "qhh.icfile.npz"), y=snap["y"], ic_yEleIS=snap["ic_yEleIS"],
                        ic_yEleSnow=snap["ic_yEleSnow"], ic_t=snap["ic_t"],
                        text=np.frombuffer(open(txtf, "rb").read(), dtype=np.uint8), _cmd=np.array(" ".join(cmd[1:])))
    sz = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"golden fixtures: {len(os.listdir(OUT))} files, {sz/1e6:.2f} MB")


if __name__ == "__main__":
    main()
