import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib
from shud_up_b200 import partition
from test_partition_gpu import _run
basin, case, nparts = sys.argv[1], sys.argv[2], int(sys.argv[3])
mesh = oracle_lib.load_case(basin, case)
Ne, Nr = int(mesh["Ne"][0]), int(mesh["Nr"][0])
mesh = dict(mesh); mesh["ele_u_satn"] = oracle_lib.oracle_prime(mesh, mesh["y"])
rhs, st, y, yd, yd_ref = _run(mesh)
with torch.cuda.stream(st):
    rhs.f_dev(0.0, y, yd); rhs.from_device_order(yd, yd_ref)
st.synchronize()
ref = yd_ref.cpu().numpy()
part = partition.assign_cells(mesh, nparts)
closures = [partition._closure_with_lakes(mesh, part, p) for p in range(nparts)]
ex = [partition.extract_cut(mesh, part, p, closures) for p in range(nparts)]
ctxs = [_run(loc) for loc, _ in ex]
_keep = []
for (loc, plan), c in zip(ex, ctxs):
    z = torch.zeros(2 * c[0].Nhalo + 3 * int(loc["n_ghost_cells"][0]) + int(loc["n_ghost_reaches"][0]) + 8, dtype=torch.float64, device="cuda")
    _keep.append(z)
    c[0].set_halo_state(z)
    c[0].f_dev(0.0, c[2], c[3], diag=True)   # allocates the diagnostic arrays before any spin-wait exists
torch.cuda.synchronize()
for (loc, plan), c in zip(ex, ctxs):
    c[0].check()
    c[0].exchange_plan_items(plan)
blobs = [ctxs[p][0].p2p_export(p) for p in range(nparts)]
for p in range(nparts):
    assert ctxs[p][0].p2p_connect_blobs(p, blobs)
torch.cuda.synchronize()
outs = []
for (loc, plan), (r, s, yy, ydd, ydr) in zip(ex, ctxs):
    r.prime(loc["y"]); r.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
    outs.append(torch.full_like(ydd, float("nan")))
torch.cuda.synchronize()
for p, (r, s, yy, ydd, ydr) in enumerate(ctxs):
    r.f_exchange_dev(0.0, yy, outs[p])
down = np.asarray(mesh["riv_down"]); seg_e = np.asarray(mesh["seg_iEle"]) - 1; seg_r = np.asarray(mesh["seg_iRiv"]) - 1
for p, ((loc, plan), (r, s, yy, ydd, ydr)) in enumerate(zip(ex, ctxs)):
    with torch.cuda.stream(s):
        r.from_device_order(outs[p], ydr)
    s.synchronize()
    got = ydr.cpu().numpy()
    nloc, nro = r.Ne, loc["_riv_ref"].size
    g, w = got[3 * nloc:3 * nloc + nro], ref[3 * Ne + loc["_riv_ref"]]
    bad = np.nonzero(g != w)[0]
    for b in bad[:6]:
        gr = loc["_riv_ref"][b]
        ups = np.nonzero(down == gr + 1)[0]
        segs = np.nonzero(seg_r == gr)[0]
        print("part", p, "local reach", b, "global", gr, "got", g[b], "want", w[b], "diff", g[b] - w[b],
              "ups", ups, "ups owner", closures[p]["riv_owner"][ups], "down", down[gr],
              "seg cells owner", part[seg_e[segs]], "code", r.check())
# ---- diag comparison for the bad partitions ----
import importlib.util
_spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py")); bench = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(bench)
for (loc, plan), (r, s, yy, ydd, ydr) in zip(ex, ctxs):
    r.prime(loc["y"]); r.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
torch.cuda.synchronize()
for p, (r, s, yy, ydd, ydr) in enumerate(ctxs):
    r.f_dev(0.0, yy, outs[p], diag=True)
torch.cuda.synchronize()
for p, ((loc, plan), (r, s, yy, ydd, ydr)) in enumerate(zip(ex, ctxs)):
    d = r.get_diag()
    ext, ne, nh = bench.extended_for_oracle(loc)
    ext["ele_u_satn"] = oracle_lib.oracle_prime(ext, ext["y"])
    o = oracle_lib.oracle_rhs(ext)
    nro = loc["_riv_ref"].size
    for name in ("QrivUp", "QrivSurf", "QrivSub", "QrivDown"):
        a, b = d[name][:nro], o[name][:nro]
        bad = np.nonzero(np.abs(a - b) > 1e-9 * (np.abs(b) + 1e-12))[0]
        if bad.size:
            print("part", p, name, "bad local reaches", bad[:5], a[bad[:5]], b[bad[:5]])
    ns = int(loc["Ns"][0])
    for name in ("QsegSurf", "QsegSub"):
        a, b = d[name], o[name][:ns]
        own_r = np.asarray(loc["seg_iRiv"]) <= nro
        bad = np.nonzero((np.abs(a - b) > 1e-9 * (np.abs(b) + 1e-12)) & own_r)[0]
        if bad.size:
            print("part", p, name, "bad segs", bad[:5], a[bad[:5]], b[bad[:5]], "cells", np.asarray(loc["seg_iEle"])[bad[:5]], "nloc", r.Ne, "ngc", int(loc["n_ghost_cells"][0]),
                  "rivs", np.asarray(loc["seg_iRiv"])[bad[:5]])
