"""The SUNDIALS-facing side of the boundary on the GPU (include/shud_sundials.h, include/shud_cvode.h):
examples/cvode_from_c.c binds it from plain C in the reference driver's order (N_VNew, SetIC2Y through the host mirror,
clones, every operation through v->ops against the flat calls bit for bit, f() as CVRhsFn, CVode with the ops-table
and the device-fused Newton-Krylov path, summary); the Python side checks f()'s checksum against the oracle and the
end state of the integration against the SAME integrator source on the host serial vector + oracle RHS."""
import os
import subprocess

import numpy as np
import pytest

import host_cv
import oracle_lib
from shud_up_b200 import api, cvode
from test_c_example import ROOT, _inputs

pytestmark = pytest.mark.gpu


def _build(tmp_path):
    exe = str(tmp_path / "cvode_from_c")
    libdir = os.path.join(ROOT, "shud_up_b200")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cvode_from_c.c"),
           "-L", libdir, "-lshud_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("basin,case", [("ccw", "rand1"), ("qhh", "rand4")])
def test_ops_table_rhs_and_integrator_from_c(tmp_path, basin, case):
    snap = oracle_lib.load_case(basin, case)
    exe = _build(tmp_path)
    mesh_path, case_path = _inputs(tmp_path, snap)
    r = subprocess.run([exe, mesh_path, case_path, "30"], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "ALL OK" in r.stdout and "FAIL" not in r.stdout, (r.stdout, r.stderr)
    # f(): the oracle primed the same way (updateforcing's satn from y), sequential sums as the C program forms them
    satn = oracle_lib.oracle_prime(snap, snap["y"])
    ref = oracle_lib.oracle_rhs(snap, u_satn=satn, want_diag=False)["ydot"]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("f: ")][0]
    s_c = float(line.split("sum(ydot)=")[1].split()[0]); sa_c = float(line.split("sum|ydot|=")[1])
    sa = float(np.sum(np.abs(ref)))
    assert abs(sa_c - sa) <= 1e-12 * sa and abs(s_c - float(np.sum(ref))) <= 1e-12 * sa
    # CVode: same integrator source, host serial vector + oracle RHS, same settings
    m = host_cv.OracleCV(snap)
    m.satn[:] = satn
    NY = int(np.asarray(snap["y"]).size)
    y = host_cv.HostVector(NY, snap["y"])
    cv = cvode.CVode(host_cv.lib(), m.f_addr, m.user_data, 0.0, y.h)
    cv.configure(rtol=1e-4, atol=1e-4, init_step=0.1, max_step=10.0)
    for tout in (10.0, 20.0, 30.0):
        cv.solve(tout, y.h)
    st = cv.stats()
    Ne = int(snap["Ne"][0])
    want = [y.array[:Ne].sum(), y.array[Ne:2 * Ne].sum(), y.array[2 * Ne:3 * Ne].sum(), y.array[3 * Ne:].sum()]
    for arm in (0, 1):
        ln = [x for x in r.stdout.splitlines() if x.startswith(f"cvode arm {arm}:")][0]
        got = [float(ln.split(k + "=")[1].split()[0]) for k in ("sum(Ysurf)", "sum(Yunsat)", "sum(Ygw)", "sum(Yriv+lake)")]
        nst = int(ln.split("nst=")[1].split()[0])
        assert abs(nst - st["nst"]) <= 1 + st["nst"] // 10, (nst, st)
        for g, w in zip(got, want):
            assert abs(g - w) <= 1e-6 * (abs(w) + 1e-3), (arm, got, want)
    cv.close(); y.close()


def test_device_vector_is_a_sundials_vector_for_python_too():
    """the same table through ctypes: clone, fill through the mirror, reductions against numpy"""
    import ctypes as C
    snap = oracle_lib.load_case("ccw", "ic")
    from shud_up_b200 import driver
    run = dict(np.load(os.path.join(oracle_lib.GOLDEN, "ccw.run.npz")))
    arm = driver.GpuArm(snap, run, fused=False)
    L = arm.lib
    x = C.c_void_p(L.N_VClone(arm.y))
    L.N_VScale(2.0, arm.y, x)
    y = np.asarray(snap["y"], dtype=np.float64)
    assert np.array_equal(arm.state_host(), y)
    hx = np.ctypeslib.as_array(L.N_VGetArrayPointer(x), shape=(arm.NY,))
    assert np.array_equal(hx, 2.0 * y)
    assert abs(L.N_VDotProd(arm.y, x) - 2.0 * float(y @ y)) <= 1e-13 * 2.0 * float(y @ y)
    assert L.N_VMaxNorm(x) == 2.0 * np.abs(y).max() and L.N_VMin(arm.y) == y.min()
    L.N_VDestroy(x)
    arm.close()
