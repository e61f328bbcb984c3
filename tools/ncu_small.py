#!/usr/bin/env python
"""A few launches of every kernel other than the cell kernel (pre-pass, river/lake, land-surface step, N_Vector
streaming + reduction) on synthetic-1M, for one ncu pass with duration + DRAM byte metrics:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv \
      --log-file gpurun_out/r01_small_kernels.csv python tools/ncu_small.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shud_up_b200 import abi, synth
from shud_up_b200.api import ShudRHS
from shud_up_b200.nvector import NVectorOps

mesh = synth.make(**synth.named("1M"))
rhs = ShudRHS(mesh)
rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
rhs.prime(mesh["y"])
st = rhs.torch_stream()
Ne = rhs.Ne
with torch.cuda.stream(st):
    y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda()
    y = torch.empty_like(y_ref); ydot = torch.empty_like(y_ref); z = torch.empty_like(y_ref)
    rhs.to_device_order(y_ref, y)
    for _ in range(3):
        for s in range(rhs.launches_per_rhs):
            rhs.f_stage_dev(s, y, ydot)
st.synchronize()
rng = np.random.default_rng(3)
tilt = rng.normal(0, 0.15, (2, Ne)); nz = 1 / np.sqrt(1 + (tilt ** 2).sum(0))
ls = {"land_nforc": [1], "land_nlc": [12], "land_nmf": [1], "land_iForc": np.ones(Ne, np.int32),
      "land_iLC": rng.integers(1, 13, Ne).astype(np.int32), "land_iMF": np.ones(Ne, np.int32),
      "land_Albedo": rng.uniform(0.1, 0.3, Ne), "land_FixPressure": rng.uniform(85, 95, Ne), "land_windH": np.full(Ne, 10.0),
      "land_nx": tilt[0] * nz, "land_ny": tilt[1] * nz, "land_nz": nz, "land_forc_z": [-9999.0], "land_gc": [1, 0, 1, 1, 1, 1],
      "land_cs": [0, 1, 0, 5.0, 0.05, 1]}
L, keep = abi.make_land(ls)
rhs.land_create(L)
rhs.land_set_state(np.zeros(Ne), np.zeros(Ne))
S = abi.ShudLandStep()
arr = {"forc": np.array([12.0, 1.5, 0.85, 2.0, 150.0]), "lai": rng.uniform(0.3, 5.0, 12), "mf": np.array([0.0013]),
       "tsr_sx": np.array([-0.6, -0.5]), "tsr_sy": np.array([-0.5, -0.4]), "tsr_sz": np.array([0.62, 0.77]),
       "tsr_wdt": np.array([18.6, 23.1])}
for k, v in arr.items():
    setattr(S, k, v.ctypes.data_as(abi._PD))
S.tsr_n, S.tsr_den, S.dt_min, S.t = 2, 41.7, 60.0, 0.0
for _ in range(3):
    rhs.land_step(S)
st.synchronize()
ops = NVectorOps(0, rhs.stream_ptr, owner=rhs)
for _ in range(3):
    ops.N_VLinearSum(1.5, y, -0.5, ydot, z)
    ops.N_VScale(2.0, y, z)
    ops.N_VDotProd(y, ydot)
    ops.N_VWrmsNorm(y, ydot)
st.synchronize()
print("ok", rhs.check())
ops.close()
