#!/usr/bin/env python
"""Newton-Krylov (BDF + SPGMR) steps on the synthetic 1M mesh: the RHS and the device N_Vector working together
the way CVODE drives them (BASELINE.json configs[3] 'RHS+SPGMR on 1 B200')."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from shud_up_b200 import synth
from shud_up_b200.api import ShudRHS
from shud_up_b200.nvector import NVectorOps
from shud_up_b200.integrator import BDFKrylov

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mesh = synth.make(**synth.named("1M"))
rhs = ShudRHS(mesh)
rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
rhs.prime(mesh["y"])
st = rhs.torch_stream()
ops = NVectorOps(0, rhs.stream_ptr, owner=rhs)
def newv():
    with torch.cuda.stream(st):
        return torch.zeros(rhs.NY, dtype=torch.float64, device="cuda")
with torch.cuda.stream(st):
    y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda()
    y = torch.empty_like(y_ref)
    rhs.to_device_order(y_ref, y)
from shud_up_b200.nvector import DeviceSPGMR
native = os.environ.get("SHUD_NATIVE_SPGMR", "1") == "1"
ls = DeviceSPGMR(ops, rhs, maxl=5) if native else None
integ = BDFKrylov(ops, newv, lambda t, a, b: rhs.f_dev(t, a, b), rhs.NY, rtol=1e-4, atol=1e-4, max_step=10.0, init_step=1e-3,
                  linear_solver=ls)
integ.init(0.0, y)
for _ in range(3):
    integ.step(1e9)
st.synchronize()
s0 = dict(integ.stats)
t0 = time.perf_counter()
for _ in range(nsteps):
    integ.step(1e9)
st.synchronize()
wall = time.perf_counter() - t0
d = {k: integ.stats[k] - s0[k] for k in s0}
code, where = rhs.check()
print(json.dumps({"bdf_steps": nsteps, "wall_ms": wall * 1e3, "stats": d, "t_sim_min": integ.t, "h_last": integ.h,
                  "rhs_calls_per_s": d["nfe"] / wall, "cell_updates_per_s": d["nfe"] * rhs.Ne / wall,
                  "ms_per_rhs_call_incl_vector_ops": wall * 1e3 / d["nfe"], "err_code": code, "native_spgmr": native}))
if ls: ls.close()
ops.close(); rhs.close()
