#!/usr/bin/env python
"""GB/s of each device N_Vector op on NY-length vectors (synthetic-1M: NY = 3,050,000), CUDA events."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from shud_up_b200.nvector import NVectorOps
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_050_000
st = torch.cuda.Stream()
ops = NVectorOps(0, st.cuda_stream)
V = [torch.randn(n, dtype=torch.float64, device="cuda") for _ in range(14)]
torch.cuda.synchronize()
def t(fn, rep=200):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(rep): fn()
    e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / rep * 1e3
x, y, z, w = V[:4]
res = {}
def rec(name, us, nbytes): res[name] = {"us": round(us, 2), "GBs": round(nbytes / us / 1e3, 1)}
rec("LinearSum", t(lambda: ops.N_VLinearSum(1.5, x, -0.5, y, z)), 24 * n)
rec("LinearSum_axpy", t(lambda: ops.N_VLinearSum(1.5, x, 1.0, y, y)), 24 * n)
rec("Scale", t(lambda: ops.N_VScale(2.0, x, z)), 16 * n)
rec("Const", t(lambda: ops.N_VConst(1.0, z)), 8 * n)
rec("Prod", t(lambda: ops.N_VProd(x, y, z)), 24 * n)
rec("DotProd(sync)", t(lambda: ops.N_VDotProd(x, y), 100), 16 * n)
rec("WrmsNorm(sync)", t(lambda: ops.N_VWrmsNorm(x, w), 100), 16 * n)
rec("MaxNorm(sync)", t(lambda: ops.N_VMaxNorm(x), 100), 8 * n)
rec("LinearCombination6", t(lambda: ops.N_VLinearCombination([1, 2, 3, 4, 5, 6], V[4:10], z)), 8 * 7 * n)
rec("ScaleAddMulti5", t(lambda: ops.N_VScaleAddMulti([1, 2, 3, 4, 5], x, V[4:9], V[9:14])), 8 * 11 * n)
rec("DotProdMulti6(sync)", t(lambda: ops.N_VDotProdMulti(x, V[4:10]), 100), 8 * 7 * n)
print(json.dumps({"n": n, "ops": res}))
