#!/usr/bin/env python
"""f() on Hilbert-range partitions that CUT the river network (partition.assign_cells / extract_cut), one process per GPU:
the whole synthetic mesh is cut into WORLD_SIZE ranges of the Hilbert curve of its cell centroids - river trees, banks and
all - and each rank runs own + ghost cells / reaches with the peer-to-peer exchange (halo pairs, ghost-cell triples,
ghost-reach stages).  Prints one JSON line: per-rank sizes, ghosts, doubles exchanged, ms per f() (max over ranks, CUDA
events), and the parity of every rank's owned entries against the CPU oracle on the same local mesh.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/cut_bench.py [nx ny] [steps]"""
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from shud_up_b200 import partition, synth  # noqa: E402
from shud_up_b200.api import ShudRHS  # noqa: E402

rank, world, lrank = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
torch.cuda.set_device(lrank)
dev = torch.device(f"cuda:{lrank}")
dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
mesh = synth.make(nx, ny, ntree=max(1, ny // 10), reaches_per_tree=nx)      # every rank builds the whole mesh
rw = float(os.environ.get("CUT_REACH_WEIGHT", "0"))
part = partition.assign_cells(mesh, world, reach_weight=rw)
closures = [partition._closure_with_lakes(mesh, part, p) for p in range(world)]
loc, plan = partition.extract_cut(mesh, part, rank, closures)
t_setup = time.time() - t0
rhs = ShudRHS(loc, device=lrank)
rhs.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
rhs.prime(loc["y"])
st = rhs.torch_stream()
with torch.cuda.stream(st):
    y_ref = torch.from_numpy(np.ascontiguousarray(loc["y"])).to(dev)
    y, yd = torch.empty_like(y_ref), torch.empty_like(y_ref)
    rhs.to_device_order(y_ref, y)
st.synchronize()
rhs.exchange_plan_items(plan)
assert rhs.p2p_connect(dist, dev), "peer mapping unavailable"
rhs.f_exchange_dev(0.0, y, yd)
with torch.cuda.stream(st):
    rhs.from_device_order(yd, y_ref)
st.synchronize()
assert rhs.check()[0] == 0
got = y_ref.cpu().numpy()
# parity of the owned entries: the CPU oracle on this rank's local mesh (halo cells as extra cells, ghosts from y)
spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
import oracle_lib  # noqa: E402
import parity  # noqa: E402
ext, ne, nh = bench.extended_for_oracle(loc)
satn = oracle_lib.oracle_prime(ext, ext["y"])
o = oracle_lib.oracle_rhs(ext, u_satn=satn, qEleE_IC=ext["qEleE_IC_in"], nthreads=max(1, (os.cpu_count() or 1) // world))
assert o["err"] == 0
sc = parity.ydot_scale(ext, o)
NE, nown, nro = ne + nh, loc["_own_ref"].size, loc["_riv_ref"].size
keep_o = np.r_[0:nown, NE:NE + nown, 2 * NE:2 * NE + nown, 3 * NE:3 * NE + nro]
keep_g = np.r_[0:nown, ne:ne + nown, 2 * ne:2 * ne + nown, 3 * ne:3 * ne + nro]
bad = parity.mismatches(got[keep_g], o["ydot"][keep_o], sc[keep_o])
ngc, ngr = int(loc["n_ghost_cells"][0]), int(loc["n_ghost_reaches"][0])
ghost_zero = bool(np.all(got[np.r_[nown:ne, ne + nown:2 * ne, 2 * ne + nown:3 * ne, 3 * ne + nro:3 * ne + rhs.Nr]] == 0.0))
for _ in range(10):
    rhs.f_exchange_dev(0.0, y, yd)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    rhs.f_exchange_dev(0.0, y, yd)
e1.record(st)
dist.barrier(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
rec = torch.tensor([ms, float(nown), float(ngc), float(ngr), float(rhs.Nhalo), float(plan["send_counts"].sum()), float(bad.size),
                    1.0 if ghost_zero else 0.0, float(nro)], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(rec) for _ in range(world)]
dist.all_gather(allr, rec)
if rank == 0:
    a = np.array([t.cpu().numpy() for t in allr])
    Ne = int(mesh["Ne"][0])
    print(json.dumps({"what": "f() on Hilbert-range partitions with cut river trees (tools/cut_bench.py)", "n_gpus": world,
                      "mesh": {"Ne": Ne, "Nr": int(mesh["Nr"][0]), "Ns": int(mesh["Ns"][0])},
                      "ms_per_f_max": float(a[:, 0].max()), "cell_updates_per_s": Ne / (float(a[:, 0].max()) * 1e-3),
                      "own_cells": a[:, 1].astype(int).tolist(), "own_reaches": a[:, 8].astype(int).tolist(),
                      "imbalance": float(a[:, 1].max() / a[:, 1].mean() - 1.0),
                      "ghost_cells": a[:, 2].astype(int).tolist(), "ghost_reaches": a[:, 3].astype(int).tolist(),
                      "halo_cells": a[:, 4].astype(int).tolist(), "doubles_sent_per_f": a[:, 5].astype(int).tolist(),
                      "reach_weight": rw, "parity_n_bad": int(a[:, 6].sum()), "ghost_ydot_zero": bool(a[:, 7].min() > 0), "setup_s": t_setup}))
rhs.close()
dist.destroy_process_group()
