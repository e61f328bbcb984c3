import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch, ctypes as C
from shud_up_b200 import synth
from shud_up_b200.api import ShudRHS, lib
mesh = synth.make(**synth.named("1M"))
rhs = ShudRHS(mesh); rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"]); rhs.prime(mesh["y"])
st = rhs.torch_stream()
with torch.cuda.stream(st):
    y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda(); y = torch.empty_like(y_ref); ydot = torch.empty_like(y_ref)
    rhs.to_device_order(y_ref, y)
st.synchronize()
def timeit(fn, n=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): fn()
    e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(json.dumps({"stage0": timeit(lambda: rhs.f_stage_dev(0, y, ydot)), "stage1": timeit(lambda: rhs.f_stage_dev(1, y, ydot)), "rhs": timeit(lambda: rhs.f_dev(0.0, y, ydot))}))
