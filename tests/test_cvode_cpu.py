"""The CVODE-shaped integrator (shud_up_b200/csrc/shud_cvode.cpp, include/shud_cvode.h) on the host serial N_Vector:
accuracy against SciPy's Radau on a stiff nonlinear system, order / step selection, CVode()'s tout / tstop semantics,
dense output, fused operations vs their loops.  SUNDIALS is absent here ("parity unpinned"): these tests pin the
restatement's behaviour, not SUNDIALS' step sequences."""
import ctypes as C

import numpy as np
import pytest
from scipy.integrate import solve_ivp

import host_cv
from shud_up_b200 import cvode

N = 40
LAM = np.linspace(0.05, 50.0, N)          # stiffness ratio 1000
KAPPA = 2.0


def _rhs(t, y):
    yl = np.concatenate(([0.0], y[:-1])); yr = np.concatenate((y[1:], [0.0]))
    return -LAM * y + KAPPA * (yl - 2.0 * y + yr) - 0.5 * y ** 3


def _make(rtol, atol, **kw):
    L = host_cv.lib()
    y = host_cv.HostVector(N, np.linspace(1.0, 2.0, N))
    par = np.concatenate(([KAPPA], LAM))
    cv = cvode.CVode(L, cvode.fn_address(L, "host_test_f"), par.ctypes.data, 0.0, y.h)
    cv.configure(rtol=rtol, atol=atol, **kw)
    cv._par = par
    return L, y, cv


@pytest.fixture(scope="module")
def truth():
    s = solve_ivp(_rhs, (0.0, 4.0), np.linspace(1.0, 2.0, N), method="Radau", rtol=1e-12, atol=1e-14, dense_output=True)
    return s.sol


@pytest.mark.parametrize("tol", [1e-4, 1e-6, 1e-8])
def test_accuracy_tracks_the_tolerance(truth, tol):
    L, y, cv = _make(tol, tol * 1e-2, max_num_steps=100000)
    assert cv.solve(4.0, y.h) == cvode.CV_SUCCESS and cv.t == 4.0
    err = np.abs(y.array - truth(4.0)) / (tol * np.abs(truth(4.0)) + tol * 1e-2)
    st = cv.stats()
    assert err.max() < 30.0, (err.max(), st)          # global error within a modest multiple of the local tolerance
    assert st["netf"] <= st["nst"] // 4 and st["ncfn"] <= st["nst"] // 4
    assert st["nni"] >= st["nst"] and st["nfe"] >= st["nni"]
    if tol <= 1e-6:
        assert st["qlast"] >= 3, st                    # the order climbs above the old integrator's limit of 2
    cv.close(); y.close()


def test_tighter_tolerance_costs_more_steps_and_higher_order():
    res = {}
    for tol in (1e-3, 1e-6, 1e-9):
        L, y, cv = _make(tol, tol, max_num_steps=100000)
        cv.solve(1.0, y.h)
        res[tol] = cv.stats()
        cv.close(); y.close()
    assert res[1e-3]["nst"] < res[1e-6]["nst"] < res[1e-9]["nst"]
    assert res[1e-9]["qlast"] >= 4


def test_normal_mode_interpolates_and_does_not_disturb_the_steps(truth):
    """CV_NORMAL returns y(tout) by dense output; the internal step sequence does not depend on where the caller
    asks for output (no stop time set; the initial step given, since CVODE's own estimate looks at the first tout)"""
    L, y1, cv1 = _make(1e-6, 1e-8, max_num_steps=100000, init_step=1e-5)
    cv1.solve(3.0, y1.h)
    L, y2, cv2 = _make(1e-6, 1e-8, max_num_steps=100000, init_step=1e-5)
    for tout in np.linspace(0.25, 3.0, 12):
        assert cv2.solve(tout, y2.h) == cvode.CV_SUCCESS and cv2.t == tout
        assert np.allclose(y2.array, truth(tout), rtol=3e-5, atol=3e-7)
    assert np.array_equal(y1.array, y2.array)
    assert cv1.stats() == cv2.stats()
    for c, v in ((cv1, y1), (cv2, y2)):
        c.close(); v.close()


def test_stop_time_is_hit_exactly_and_reported():
    L, y, cv = _make(1e-5, 1e-7, max_num_steps=100000)
    cv.set_stop_time(0.7)
    assert cv.solve(2.0, y.h) == cvode.CV_TSTOP_RETURN and cv.t == 0.7
    assert cv.stats()["tcur"] == 0.7                   # the step was clipped to land on tstop
    assert cv.solve(2.0, y.h) == cvode.CV_SUCCESS and cv.t == 2.0
    cv.close(); y.close()


def test_one_step_mode_and_dky(truth):
    L, y, cv = _make(1e-6, 1e-8, init_step=1e-4)
    ts = []
    for _ in range(25):
        assert cv.solve(10.0, y.h, itask=cvode.CV_ONE_STEP) == cvode.CV_SUCCESS
        ts.append(cv.t)
    st = cv.stats()
    assert st["nst"] == 25 and st["hinused"] == 1e-4 and np.all(np.diff(ts) > 0)
    mid = host_cv.HostVector(N); d1 = host_cv.HostVector(N)
    tm = cv.t - 0.5 * st["hlast"]
    cv.get_dky(tm, 0, mid.h); cv.get_dky(tm, 1, d1.h)
    assert np.allclose(mid.array, truth(tm), rtol=1e-4, atol=1e-6)
    assert np.allclose(d1.array, _rhs(tm, truth(tm)), rtol=2e-2, atol=1e-4)
    with pytest.raises(cvode.CVError):
        cv.get_dky(cv.t + 1.0, 0, mid.h)               # CV_BAD_T outside the last step
    for v in (y, mid, d1):
        v.close()
    cv.close()


def test_step_limits_and_work_limit():
    L, y, cv = _make(1e-4, 1e-6, max_step=0.01, max_num_steps=1000000)
    cv.solve(1.0, y.h)
    st = cv.stats()
    assert st["hlast"] <= 0.01 * (1 + 1e-12) and st["nst"] >= 100
    cv.close(); y.close()
    L, y, cv = _make(1e-8, 1e-10, max_num_steps=20)
    with pytest.raises(cvode.CVError) as e:
        cv.solve(4.0, y.h)
    assert e.value.flag == -1 and cv.stats()["nst"] == 20     # CV_TOO_MUCH_WORK after mxstep steps, t = tn
    cv.close(); y.close()


def test_fused_table_members_equal_their_loops():
    """N_VLinearCombination / N_VScaleAddMulti / N_VDotProdMulti absent -> the generic layer loops over the standard
    operations: same bits, same steps"""
    out = []
    for fused in (True, False):
        L, y, cv = _make(1e-6, 1e-8, max_num_steps=100000)
        if not fused:
            L.N_VDisableFused_HostSerial(y.h)
            cv.close()
            par = np.concatenate(([KAPPA], LAM))
            y.array[:] = np.linspace(1.0, 2.0, N)
            cv = cvode.CVode(L, cvode.fn_address(L, "host_test_f"), par.ctypes.data, 0.0, y.h)   # clones inherit the table
            cv.configure(rtol=1e-6, atol=1e-8, max_num_steps=100000)
            cv._par = par
        cv.solve(2.0, y.h)
        out.append((y.array.copy(), cv.stats()))
        cv.close(); y.close()
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


def test_oracle_rhs_under_the_integrator_converges_with_the_tolerance():
    """ccw, 6 hours from the initial condition under constant forcing: the run at rtol = atol = 1e-4 (the reference's
    setting) stays within a few local tolerances of the run at 1e-7"""
    import oracle_lib
    snap = oracle_lib.load_case("ccw", "ic")
    NY = int(np.asarray(snap["y"]).size)
    ends = {}
    for tol in (1e-4, 1e-7):
        m = host_cv.OracleCV(snap)
        y = host_cv.HostVector(NY, snap["y"])
        cv = cvode.CVode(host_cv.lib(), m.f_addr, m.user_data, 0.0, y.h)
        cv.configure(rtol=tol, atol=tol, init_step=1.0, max_step=10.0)
        t = 0.0
        while t < 360.0:
            t += 10.0
            cv.solve(t, y.h)
        ends[tol] = (y.array.copy(), cv.stats())
        cv.close(); y.close()
    ya, yb = ends[1e-4][0], ends[1e-7][0]
    wrms = np.sqrt(np.mean(((ya - yb) / (1e-4 * np.abs(yb) + 1e-4)) ** 2))
    assert wrms < 1.0, (wrms, ends[1e-4][1], ends[1e-7][1])
    assert ends[1e-4][1]["nst"] >= 36 and ends[1e-7][1]["nst"] > ends[1e-4][1]["nst"]


def _run_steps(hooked, nsteps, rhs="test"):
    L = host_cv.lib()
    if rhs == "test":
        L_, y, cv = _make(1e-6, 1e-8, max_num_steps=100000)
        keep = None
    else:
        import oracle_lib
        snap = oracle_lib.load_case("ccw", "ic")
        keep = host_cv.OracleCV(snap)
        keep.satn[:] = oracle_lib.oracle_prime(snap, snap["y"])
        y = host_cv.HostVector(snap["y"].size, np.ascontiguousarray(snap["y"]))
        cv = cvode.CVode(L, keep.f_addr, keep.user_data, 0.0, y.h)
        cv.configure(rtol=1e-4, atol=1e-4, init_step=1e-3, max_step=10.0)
    fz, calls = None, None
    if hooked:
        fz = cvode.Fused()
        L.host_cv_fused_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(cvode.Fused)]
        L.host_cv_fused_calls.argtypes = [C.POINTER(cvode.Fused), C.POINTER(C.c_long)]
        L.host_cv_fused_destroy.argtypes = [C.POINTER(cvode.Fused)]
        assert L.host_cv_fused_create(cv._h, y.h, C.byref(fz)) == 0
        cv.set_fused(fz)
    traj = []
    for _ in range(nsteps):
        cv.solve(1e9, y.h, itask=cvode.CV_ONE_STEP)
        traj.append((cv.t, y.array.copy()))
    st = cv.stats()
    if hooked:
        c4 = (C.c_long * 4)()
        L.host_cv_fused_calls(C.byref(fz), c4)
        calls = list(c4)
    cv.close()
    if hooked:
        L.host_cv_fused_destroy(C.byref(fz))
    y.close()
    return traj, st, calls


@pytest.mark.parametrize("rhs,nsteps", [("test", 150), ("oracle", 40)])
def test_hooked_newton_step_and_predictor_reproduce_the_plain_integrator(rhs, nsteps):
    """shud_cv_fused.predict / newton_step / ewt_set_norm (single fused kernels on the GPU) restated with the generic
    vector operations (oracle/host_cv.c): the hooked control flow of the integrator - predictor that also primes the
    Newton iteration, one hook per Newton iteration, weights + norm in one call - takes the same steps and produces the
    same vectors, bit for bit, as the plain route (cvPredict / cvNls / cvLsSolve / cvEwtSet sequence)"""
    plain, st0, _ = _run_steps(False, nsteps, rhs)
    hook, st1, calls = _run_steps(True, nsteps, rhs)
    # predictor and Newton hook every step; the completion hook too, and the weights + norm it leaves are the ones the
    # next step starts with (ewt_set_norm is left with nothing to do after the first step)
    assert calls[0] >= nsteps and calls[1] >= nsteps and calls[3] == nsteps and calls[2] <= 1, calls
    for k in ("nst", "nfe", "nfeLS", "nni", "nli", "ncfn", "netf", "ncfl", "qlast", "hlast"):
        assert st0[k] == st1[k], (k, st0, st1)
    for (t0, y0), (t1, y1) in zip(plain, hook):
        assert t0 == t1
        assert np.array_equal(y0, y1)
