#!/usr/bin/env python
"""Launch the cell kernel a few times for an ncu capture (ncu -k regex:k_fused -s 3 -c 1 ... python tools/ncu_stage.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shud_up_b200 import synth
from shud_up_b200.api import ShudRHS
mesh = synth.make(**synth.named(sys.argv[1] if len(sys.argv) > 1 else "1M"))
rhs = ShudRHS(mesh)
rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
rhs.prime(mesh["y"])
st = rhs.torch_stream()
with torch.cuda.stream(st):
    y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda()
    y = torch.empty_like(y_ref); ydot = torch.empty_like(y_ref)
    rhs.to_device_order(y_ref, y)
    stages = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(rhs.launches_per_rhs))
    for _ in range(6):
        for s in stages:
            rhs.f_stage_dev(s, y, ydot)
st.synchronize()
print("ok", rhs.check())
