import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import oracle_lib
from shud_up_b200.api import ShudRHS
basin, case = sys.argv[1], sys.argv[2]
mode = sys.argv[3]
snap = oracle_lib.load_case(basin, case)
rhs = ShudRHS(snap)
rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
rhs.set_carried(snap["ele_u_satn"])
y = np.ascontiguousarray(snap["y"])
st = rhs.torch_stream()
if mode == "host":
    ydot = np.full_like(y, np.nan)
    print("rc", rhs.f(0.0, y, ydot))
else:
    with torch.cuda.stream(st):
        yr = torch.from_numpy(y).cuda(); yd = torch.empty_like(yr); ydd = torch.empty_like(yr)
        rhs.to_device_order(yr, yd)
        if mode == "stage":
            for s_ in (1, 2, 3):
                rhs.f_stage_dev(s_, yd, ydd); st.synchronize(); print("stage", s_, "ok", rhs.check())
        else:
            rhs.f_dev(0.0, yd, ydd, diag=(mode == "diag"))
    st.synchronize()
    print("check", rhs.check())
print("done", mode)
