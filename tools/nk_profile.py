#!/usr/bin/env python
"""The Newton-Krylov step of bench.py's newton_krylov block (C integrator, device-fused pieces, synthetic-1M) on its own:
wall time per BDF step and per RHS call; run it under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/nk_launches.csv python tools/nk_profile.py 6
for the launch list of the same steps (tools/nk_launches.py sums it per kernel)."""
import os, sys, time, json
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shud_up_b200 import synth, cvode as _cv
from shud_up_b200.api import ShudRHS, lib as _lib

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mesh = synth.make(**synth.named(sys.argv[2] if len(sys.argv) > 2 else "1M"))
rhs = ShudRHS(mesh)
rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
rhs.prime(mesh["y"])
L = _cv.bind(_lib())
L.N_VNew_ShudB200.restype = C.c_void_p
L.N_VNew_ShudB200.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
L.N_VCopyToDevice_ShudB200.argtypes = [C.c_void_p]
ws = C.c_void_p()
assert L.shud_nv_ws_create(0, C.c_void_p(rhs.stream_ptr), C.byref(ws)) == 0
yv = C.c_void_p(L.N_VNew_ShudB200(rhs.NY, ws, rhs._h, None))
np.ctypeslib.as_array(L.N_VGetArrayPointer(yv), shape=(rhs.NY,))[:] = mesh["y"]
assert L.N_VCopyToDevice_ShudB200(yv) == 0
cvi = _cv.CVode(L, _cv.fn_address(L, "shud_b200_f"), rhs._h.value, 0.0, yv)
cvi.configure(rtol=1e-4, atol=1e-4, init_step=1e-3, max_step=10.0)
fz = _cv.Fused()
L.shud_b200_cv_fused_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_cv.Fused)]
L.shud_b200_cv_fused_destroy.argtypes = [C.POINTER(_cv.Fused)]
if os.environ.get("NK_FUSED", "1") != "0":
    assert L.shud_b200_cv_fused_create(rhs._h, ws, 5, C.byref(fz)) == 0
    cvi.set_fused(fz)
for _ in range(int(os.environ.get("NK_WARM", "3"))):
    cvi.solve(1e9, yv, itask=_cv.CV_ONE_STEP)
torch.cuda.synchronize()
s0 = cvi.stats()
t0 = time.perf_counter()
for _ in range(nsteps):
    cvi.solve(1e9, yv, itask=_cv.CV_ONE_STEP)
torch.cuda.synchronize()
w = time.perf_counter() - t0
s1 = cvi.stats()
nrhs = (s1["nfe"] + s1["nfeLS"]) - (s0["nfe"] + s0["nfeLS"])
print(json.dumps({"steps": nsteps, "rhs_calls": nrhs, "nni": s1["nni"] - s0["nni"], "nli": s1["nli"] - s0["nli"],
                  "order": s1["qlast"], "ms_per_step": w * 1e3 / nsteps, "ms_per_rhs_call": w * 1e3 / max(nrhs, 1),
                  "y_sum": float(np.ctypeslib.as_array(L.N_VGetArrayPointer(yv), shape=(rhs.NY,)).sum())}))
