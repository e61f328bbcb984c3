"""The boundary from plain C (examples/rhs_from_c.c): compiles against include/shud_b200.h with gcc and links the
library; without a device it exits with the 'no CUDA device' code (no CPU fallback), on the GPU box its checksum of
ydot equals the one of the Python binding on the same inputs."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "rhs_from_c")
    libdir = os.path.join(ROOT, "shud_up_b200")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "rhs_from_c.c"),
           "-L", libdir, "-lshud_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _inputs(tmp_path, snap):
    mesh_path, case_path = str(tmp_path / "m.shudb200"), str(tmp_path / "case.bin")
    api.mesh_save(mesh_path, snap)
    Ne = int(snap["Ne"][0])
    with open(case_path, "wb") as fp:
        np.ascontiguousarray(snap["y"], dtype=np.float64).tofile(fp)
        for k in ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "fu_Surf", "fu_Sub", "qElePrep", "qEleE_IC_in"):
            a = np.ascontiguousarray(snap[k], dtype=np.float64)
            assert a.size == Ne, k
            a.tofile(fp)
    return mesh_path, case_path


def test_c_example_builds_and_refuses_to_run_without_a_device(tmp_path):
    import torch
    exe = _build(tmp_path)
    mesh_path, case_path = _inputs(tmp_path, oracle_lib.load_case("ccw", "rand1"))
    r = subprocess.run([exe, mesh_path, case_path], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_example_matches_the_python_binding(tmp_path):
    import torch
    snap = oracle_lib.load_case("qhh", "rand4")
    exe = _build(tmp_path)
    mesh_path, case_path = _inputs(tmp_path, snap)
    r = subprocess.run([exe, mesh_path, case_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rhs = api.ShudRHS(snap)
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    rhs.prime(snap["y"])
    y = torch.from_numpy(np.ascontiguousarray(snap["y"])).pin_memory()
    yd = torch.empty_like(y).pin_memory()
    rhs.f(0.0, y, yd)
    v = yd.numpy()
    s, sa = 0.0, 0.0
    for x in v:  # the C program sums sequentially
        s += x; sa += abs(x)
    out = r.stdout.strip()
    assert f"sum(ydot)={s:.17g}" in out and f"sum|ydot|={sa:.17g}" in out, (out, s, sa)
