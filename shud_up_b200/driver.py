"""The reference's outer time loop (src/Model/shud.cpp:91-155) around the GPU path: per forcing step upload
what updateforcing()+ET() produced, then let the integrator carry y to the next output time in SolverStep
chunks, and sample the outlet discharge from a diagnostic RHS at the accepted solution
(the reference does the same before sampling its water balance, shud.cpp:138-149)."""
import time

import numpy as np

from .integrator import BDFKrylov


def run(model, fseq, y0, n_steps, solver_step=10.0, et_step=60.0, rtol=1e-4, atol=1e-4, max_step=10.0, init_step=1.0,
        max_order=2, outlet_reaches=None):
    """model: adapter with  ops, new_vector(), rhs(t,y,ydot), set_forcing(k), load_state(y0)->device vector,
    state_to_host(y)->numpy (reference order), outlet_flux(t, y)->numpy QrivDown at the outlets.
    fseq: dict of [n][Ne] forcing arrays; returns dict(t, y_end, q_out[n_steps][n_outlets], stats, wall_s)."""
    y = model.load_state(y0)
    integ = BDFKrylov(model.ops, model.new_vector, model.rhs, int(y0.size), rtol=rtol, atol=atol, max_step=max_step,
                      init_step=init_step, max_order=max_order, linear_solver=getattr(model, "linear_solver", None))
    t = float(fseq["fseq_t"][0])
    integ.init(t, y)
    q_out, times = [], []
    t0 = time.perf_counter()
    for k in range(n_steps):
        model.set_forcing(k)
        tend = t + et_step
        while t < tend - 1e-9:
            tnext = min(t + solver_step, tend)
            ycur = integ.advance(tnext)
            t = tnext
        q_out.append(model.outlet_flux(t, ycur))
        times.append(t)
    wall = time.perf_counter() - t0
    return dict(t=np.array(times), y_end=model.state_to_host(ycur), q_out=np.array(q_out), stats=dict(integ.stats),
                wall_s=wall, sim_days_per_wall_s=(times[-1] - float(fseq["fseq_t"][0])) / 1440.0 / wall)


class GpuModel:
    """adapter: ShudRHS + device N_Vector (everything stays on the device, device order)"""

    def __init__(self, mesh, fseq, device=0, land=None):
        """land: a --land-seq snapshot (abi.make_land / abi.land_steps inputs): the land-surface step then runs on
        the device (shud_b200_land_step) instead of uploading the forcing arrays of `fseq` every ET step"""
        import torch
        from .api import ShudRHS
        from .nvector import NVectorOps
        self.torch, self.mesh, self.fseq = torch, mesh, fseq
        self.shud = ShudRHS(mesh, device=device)
        self.stream = self.shud.torch_stream()
        self.ops = NVectorOps(device, self.shud.stream_ptr, owner=self.shud)
        from .nvector import DeviceSPGMR
        self.linear_solver = DeviceSPGMR(self.ops, self.shud, maxl=5)
        self.dev = torch.device(f"cuda:{device}")
        self.Ne, self.Nr = self.shud.Ne, self.shud.Nr
        self.outlets = np.nonzero(np.asarray(mesh["riv_down"]) < 0)[0]
        self._scratch = self.new_vector()
        self.land_steps = None
        if land is not None:
            from . import abi
            L, self._land_keep = abi.make_land(land)
            self.shud.set_forcing(mesh, qEleE_IC=np.zeros(self.Ne))  # BC arrays once; the rest is rewritten per step
            self.shud.land_create(L)
            self.shud.land_set_state(land["land_yEleSnow0"], land["land_yEleIS0"])
            self.land_steps = [(S, keep) for _, S, keep in abi.land_steps(land)]

    def close(self):
        self.torch.cuda.synchronize()
        self.linear_solver.close()
        self.ops.close()
        self.shud.close()

    def new_vector(self):
        with self.torch.cuda.stream(self.stream):
            return self.torch.zeros(self.shud.NY, dtype=self.torch.float64, device=self.dev)

    def rhs(self, t, y, ydot):
        self.shud.f_dev(t, y, ydot)

    def set_forcing(self, k):
        if self.land_steps is not None:
            self.shud.land_step(self.land_steps[k][0])
            return
        f = {n: self.fseq["fseq_" + n][k] for n in ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "qElePrep")}
        f["fu_Surf"] = np.ones(self.Ne); f["fu_Sub"] = np.ones(self.Ne)
        self.shud.set_forcing(f, qEleE_IC=self.fseq["fseq_qEleE_IC"][k])

    def load_state(self, y0):
        self.shud.set_carried(np.zeros(self.Ne))     # what the reference's first updateforcing() sees
        with self.torch.cuda.stream(self.stream):
            yr = self.torch.from_numpy(np.ascontiguousarray(y0)).to(self.dev)
            y = self.torch.empty_like(yr)
            self.shud.to_device_order(yr, y)
        self.stream.synchronize()
        return y

    def state_to_host(self, y):
        with self.torch.cuda.stream(self.stream):
            out = self.torch.empty_like(y)
            self.shud.from_device_order(y, out)
        self.stream.synchronize()
        return out.cpu().numpy()

    def outlet_flux(self, t, y):
        """QrivDown at the outlet reaches, from a diagnostic RHS at the accepted solution.  Like the reference's
        own direct f() calls (src/Model/shud.cpp:127,141) this moves the carried state on."""
        self.shud.f_dev(t, y, self._scratch, diag=True)
        code, where = self.shud.check()
        if code:
            raise RuntimeError(f"RHS error code {code} at {where}")
        return self.shud.get_diag()["QrivDown"][self.outlets].copy()


# ------------------------------------------------------------------------------------------------------------------
# The same outer loop on the library's CVODE-shaped integrator (include/shud_cvode.h, csrc/shud_cvode.cpp): the time
# loop below only sequences land-surface step -> CVode(tnext) -> sampling, as src/Model/shud.cpp:91-155 does; the
# Newton-Krylov iteration and every vector operation run in C on SUNDIALS-layout N_Vectors.
# ------------------------------------------------------------------------------------------------------------------
def river_storage(mesh, yriv):
    """WaterBalanceDiag::basinRiverStorage_m3 (WaterBalanceDiag.cpp:183-196): sum of u_CSarea * Length"""
    y = np.asarray(yriv, dtype=np.float64)
    w0, s, ln = (np.asarray(mesh["riv_" + n], dtype=np.float64) for n in ("BottomWidth", "bankslope", "Length"))
    cs = np.maximum(y * (w0 + y * s), 0.0)          # fun_CrossArea, clamped like updateRiver (River.cpp:49-62)
    return float(np.sum(cs * ln))


class BasinBudget:
    """the 9-column basin budget of WaterBalanceDiag (src/Model/WaterBalanceDiag.cpp:401-636) on SolverStep samples:
    storage change of cells (surface + Sy (unsat + gw) + snow + canopy) and reaches against precipitation, evaporation
    and outlet discharge integrated from the rates at the accepted solution (backward Euler, the reference's default)."""

    def __init__(self, mesh):
        self.mesh = mesh
        from . import abi
        self.area = np.asarray(mesh[abi._key(mesh, "area")], dtype=np.float64)
        self.sy = np.asarray(mesh[abi._key(mesh, "Sy")], dtype=np.float64)
        self.Ne, self.Nr = self.area.size, np.asarray(mesh["riv_down"]).size
        self.outlets = np.nonzero(np.asarray(mesh["riv_down"]) < 0)[0]
        self.p = self.et = self.qout = 0.0
        self.s0 = None
        self.last_t = None

    def storage(self, y, snow, ics):
        Ne = self.Ne
        depth = y[:Ne] + self.sy * y[Ne:2 * Ne] + self.sy * y[2 * Ne:3 * Ne] + snow + ics
        return float(np.sum(depth * self.area)) + river_storage(self.mesh, y[3 * Ne:3 * Ne + self.Nr])

    def sample(self, t, y, snow, ics, prcp, ic_raw, diag):
        s = self.storage(y, snow, ics)
        if self.s0 is None:
            self.s0, self.s_last, self.last_t = s, s, t
            return
        dt = t - self.last_t
        et3 = diag["qEs"] + diag["qEu"] + diag["qEg"] + diag["qTu"] + diag["qTg"]
        self.p += float(np.sum(prcp * self.area)) * dt
        self.et += float(np.sum((ic_raw + et3) * self.area)) * dt
        self.qout += float(np.sum(diag["QrivDown"][self.outlets])) * dt
        self.s_last, self.last_t = s, t

    def result(self):
        ds = self.s_last - self.s0
        net = self.p - self.et - self.qout
        return dict(dS_m3=ds, P_m3=self.p, ET_m3=self.et, Qout_m3=self.qout, resid_m3=ds - net,
                    resid_rel=(ds - net) / max(abs(self.p), abs(self.qout), abs(ds), 1e-300))


def run_cv(arm, run, n_steps=None, sample_every=1):
    """arm: GpuArm below, or the checker arm of tests/host_cv.py (same interface).  run: a tests/golden/<basin>.run.npz
    snapshot (tools/make_golden_runs.py): per-SolverStep land-surface inputs + run_cfg = [rtol, atol, init step,
    SolverStep = MaxStep, t0, steps].  Mirrors shud.cpp:91-155 with ETStep >= SolverStep: per SolverStep one
    land-surface step (dt = SolverStep), CVode(tnext, CV_NORMAL), summary(), a direct f() at the accepted solution
    (what the water-balance sampler does, shud.cpp:138-141), sampling of the outlet discharge."""
    from . import cvode
    rtol, atol, h0, dt, t0, n_all = [float(v) for v in run["run_cfg"]]
    n_steps = int(n_all) if n_steps is None else int(n_steps)
    cv = cvode.CVode(arm.lib, arm.f_addr, arm.user_data, t0, arm.y)
    cv.configure(rtol=rtol, atol=atol, init_step=h0, max_step=dt, min_step=1e-6, max_num_steps=1000000, maxl=0)
    if getattr(arm, "fused", None) is not None:
        cv.set_fused(arm.fused)
    budget = BasinBudget(arm.mesh)
    q_out, times = [], []
    t = t0
    wall0 = time.perf_counter()
    for k in range(n_steps):
        arm.land_step(k)
        tnext = t0 + (k + 1) * dt
        cv.solve(tnext, arm.y)
        t = tnext
        if (k + 1) % sample_every == 0 or k == n_steps - 1:
            yh = arm.state_host()
            d = arm.diag(t)
            lo = arm.land_out()
            budget.sample(t, yh, lo["yEleSnow"], lo["yEleIS"], lo["qElePrep"], lo["qEleE_IC"], d)
            q_out.append(d["QrivDown"][budget.outlets].copy())
            times.append(t)
    wall = time.perf_counter() - wall0
    st = cv.stats()
    cv.close()
    return dict(t=np.array(times), q_out=np.array(q_out), y_end=arm.state_host(), stats=st, wall_s=wall,
                sim_days_per_wall_s=(t - t0) / 1440.0 / wall, budget=budget.result())


class GpuArm:
    """the device arm: ShudRHS + device land-surface step + SHUD B200 N_Vector + shud_b200_f, optionally with the
    device-fused Newton-Krylov pieces (shud_b200_cv_fused_create)"""

    def __init__(self, mesh, run, device=0, fused=True):
        import ctypes as C

        import torch

        from . import abi, cvode
        from .api import ShudRHS, lib
        self.torch, self.mesh = torch, mesh
        self.lib = cvode.bind(lib())
        L = self.lib
        self.shud = ShudRHS(mesh, device=device)
        self.Ne, self.Nr, self.NY = self.shud.Ne, self.shud.Nr, self.shud.NY
        ws = C.c_void_p()
        rc = L.shud_nv_ws_create(int(device), C.c_void_p(self.shud.stream_ptr), C.byref(ws))
        if rc:
            raise RuntimeError(f"shud_nv_ws_create: {rc}")
        self.ws = ws
        L.N_VNew_ShudB200.restype = C.c_void_p
        L.N_VNew_ShudB200.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.N_VCopyToDevice_ShudB200.argtypes = [C.c_void_p]
        L.N_VSummary_ShudB200.restype = C.POINTER(C.c_double)
        L.N_VSummary_ShudB200.argtypes = [C.c_void_p]
        L.N_VGetDeviceArrayPointer_ShudB200.restype = C.c_void_p
        L.N_VGetDeviceArrayPointer_ShudB200.argtypes = [C.c_void_p]
        self.y = C.c_void_p(L.N_VNew_ShudB200(self.NY, ws, self.shud._h, None))
        self.ydot = C.c_void_p(L.N_VClone(self.y))
        # SetIC2Y (MD_initialize.cpp:117-135): host writes through the array pointer, then one push to the device
        host = np.ctypeslib.as_array(L.N_VGetArrayPointer(self.y), shape=(self.NY,))
        host[:] = np.asarray(mesh["y"], dtype=np.float64)
        if L.N_VCopyToDevice_ShudB200(self.y):
            raise RuntimeError("N_VCopyToDevice_ShudB200")
        self.f_addr = cvode.fn_address(L, "shud_b200_f")
        self.user_data = self.shud._h.value
        self.fused = None
        if fused:
            self.fused = cvode.Fused()
            L.shud_b200_cv_fused_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(cvode.Fused)]
            L.shud_b200_cv_fused_destroy.argtypes = [C.POINTER(cvode.Fused)]
            rc = L.shud_b200_cv_fused_create(self.shud._h, ws, 5, C.byref(self.fused))
            if rc:
                raise RuntimeError(f"shud_b200_cv_fused_create: {rc}")
        land, self._land_keep = abi.make_land(run)
        self.shud.set_forcing(mesh, qEleE_IC=np.zeros(self.Ne))   # BC arrays once; the land step rewrites the rest
        self.shud.set_carried(np.zeros(self.Ne))                   # what the reference's first updateforcing() sees
        self.shud.land_create(land)
        self.shud.land_set_state(run["land_yEleSnow0"], run["land_yEleIS0"])
        self._steps = abi.land_steps(run)
        self._step_k = -1
        self._cur = None

    def land_step(self, k):
        while self._step_k < k:
            self._step_k, S, keep = next(self._steps)
            self._cur = (S, keep)
        self.shud.land_step(self._cur[0])
        self._lo = None

    def land_out(self):
        """the land step's outputs as the host would read them after ET() (qEleE_IC before any f() clipped it is not
        kept on the device: the RHS rewrites it in place, as the reference's does - the budget uses the value the
        accepted-solution f() left, which equals the raw one whenever the canopy store covers the evaporation)"""
        return self.shud.land_get()

    def state_host(self):
        p = self.lib.N_VSummary_ShudB200(self.y)
        return np.ctypeslib.as_array(p, shape=(self.NY,)).copy()

    def diag(self, t):
        import ctypes as C
        yd = self.lib.N_VGetDeviceArrayPointer_ShudB200(self.y)
        dd = self.lib.N_VGetDeviceArrayPointer_ShudB200(self.ydot)
        from .api import _chk, lib
        _chk(lib().shud_b200_rhs_diag_dev(self.shud._h, float(t), C.c_void_p(yd), C.c_void_p(dd)), "rhs_diag_dev")
        code, where = self.shud.check()
        if code:
            raise RuntimeError(f"RHS error code {code} at {where}")
        return self.shud.get_diag()

    def close(self):
        import ctypes as C
        if self.fused is not None:
            self.lib.shud_b200_cv_fused_destroy(C.byref(self.fused))
            self.fused = None
        self.lib.N_VDestroy(self.ydot); self.lib.N_VDestroy(self.y)
        self.lib.shud_nv_ws_destroy(self.ws)
        self.shud.close()
