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
