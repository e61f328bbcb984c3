#!/usr/bin/env python
"""Per-kernel CUDA-event timing of the RHS on the synthetic 1M mesh (developer loop; bench.py is the record)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shud_up_b200 import synth
from shud_up_b200.api import ShudRHS

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
size = sys.argv[2] if len(sys.argv) > 2 else "1M"
mesh = synth.make(**synth.named(size))
order = os.environ.get("QB_ORDER", "hilbert")  # experiment: what the locality ordering is worth
if order == "xsort": mesh["ele_y"] = np.zeros_like(mesh["ele_y"])
elif order == "ysort": mesh["ele_x"] = np.zeros_like(mesh["ele_x"])
elif order == "none": mesh.pop("ele_x"); mesh.pop("ele_y")
rhs = ShudRHS(mesh)
rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
rhs.prime(mesh["y"])
st = rhs.torch_stream()
with torch.cuda.stream(st):
    y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda()
    y = torch.empty_like(y_ref); ydot = torch.empty_like(y_ref)
    rhs.to_device_order(y_ref, y)
st.synchronize()
def timeit(fn, n):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): fn()
    e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
res = {"rhs_us": timeit(lambda: rhs.f_dev(0.0, y, ydot), steps)}
for s in range(rhs.launches_per_rhs):
    res[f"stage{s}_us"] = timeit(lambda: rhs.f_stage_dev(s, y, ydot), steps)
Ne, Nr, Ns = rhs.Ne, rhs.Nr, rhs.Ns
b = 392 * Ne + 124 * Nr + 72 * Ns
res["rhs_GBs"] = b / res["rhs_us"] / 1e3
res["Gcells_s"] = Ne / res["rhs_us"] / 1e3
with torch.cuda.stream(st):
    rhs.prime(mesh["y"]); rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
    rhs.f_dev(0.0, y, ydot); rhs.from_device_order(ydot, y_ref)
st.synchronize()
yd = y_ref.cpu().numpy()
res["sum"] = float(np.cumsum(yd)[-1]); res["asum"] = float(np.cumsum(np.abs(yd))[-1])
res["code"] = rhs.check()[0]
print(json.dumps(res))
