#!/usr/bin/env python
"""bench.py - RHS cell-updates/s of the SHUD hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            the CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference's algorithm on the host cores

One "step" = one complete f(t,y,ydot) over the synthetic mesh (every cell, segment, reach, lake).
N=1 workload: BASELINE.json configs[3], the synthetic 1M-triangle mesh + 50k reaches + 150k segments
(seed 20240611).  N>1: weak scaling, one 1M-cell horizontal stripe of the N x 1M-cell
mesh per rank (configs[4] at N=8 is the 8M-cell mesh); stripes follow the river-tree bands, so only
cell states cross the cut.
Inputs are 400 MB per rank (> the 126 MB L2), so no L2 flush is needed between timed steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rhs_cell_updates_per_sec"
UNIT = "cell-updates/s"
B_CELL, B_RIV, B_SEG = 392, 124, 72  # algorithmic bytes per unit and f() call


WORKLOAD = "synthetic-1M per GPU: 1,000,000 cells / 50,000 reaches / 150,000 segments, one f() per step"


def config_for(world):
    """the workload description both arms print (identical dicts: the driver compares them)"""
    return {"workload": WORKLOAD, "seed": 20240611, "l2": "inputs 400 MB per rank > 126 MB L2, no flush needed",
            "multi_gpu": (f"{world} stripes of 1M cells of the {world}M-cell mesh, per-f() halo exchange of the boundary "
                          "cells' (Ysurf, Ygw)") if world > 1 else "single GPU"}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 100 ms; nvidia-smi, the
    recipe's command, once as a cross-check - spawning it takes ~1 s on these hosts, too slow to sample with)."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.th = threading.Thread(target=self.run, daemon=True)
        self.smi = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop:
                self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx, int(get_reasons(h))))
                time.sleep(0.1)
        except Exception:
            pass
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            self.smi = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader"],
                                      capture_output=True, text=True, timeout=10).stdout.strip()
        except Exception:
            pass

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=15)

    def summary(self):
        # NVML reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for bit, n in names.items() if r[2] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": reasons, "samples": len(sm), "nvidia_smi_after": self.smi}


def stripe_mesh(world, rank):
    """the rank's share of the N x 1M-cell mesh.  N=1: synthetic-1M (1000 x 500 quads).  N>1: the mesh of
    1000 x 500N quads (N=8: 8,000,000 cells / 400,000 reaches / 1,200,000 segments, configs[4]) cut into N
    horizontal stripes of 500 quad rows = 1M cells each; a rank builds only its own stripe plus its halo."""
    from shud_up_b200 import synth
    cfg = synth.named("1M")
    if world == 1:
        return synth.make(cfg["nx"], cfg["ny"], ntree=cfg["ntree"], reaches_per_tree=cfg["reaches_per_tree"])
    rows = cfg["ny"]
    return synth.make(cfg["nx"], rows * world, ntree=cfg["ntree"] * world, reaches_per_tree=cfg["reaches_per_tree"],
                      rows=(rank * rows, (rank + 1) * rows), stripe_rows=rows)


def cpu_oracle_time(mesh, nthreads, budget_s=15.0, max_calls=50, warm_calls=1, want_ydot=False):
    """time the CPU restatement of the reference f() (oracle/, the checker) on the host cores"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import oracle_lib
    from shud_up_b200 import abi
    L = oracle_lib.lib()
    ms, keep = abi.make_mesh(mesh)
    satn = oracle_lib.oracle_prime(mesh, mesh["y"])
    eic = np.array(mesh["qEleE_IC_in"], dtype=np.float64, copy=True)
    fs, keep2 = abi.make_forcing(mesh, qEleE_IC=eic)
    y = np.ascontiguousarray(mesh["y"], dtype=np.float64)
    ydot = np.empty_like(y)
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    call = lambda: L.shud_oracle_rhs(C.byref(ms), C.byref(fs), pd(satn), pd(eic), pd(y), pd(ydot), None, nthreads)
    ydot0 = None
    for k in range(max(1, warm_calls)):  # warm-up (first touch of the workspace)
        call()
        if k == 0 and want_ydot:  # the first call after priming: what the parity block compares the GPU with
            ydot0 = ydot.copy()
    n, t0 = 0, time.perf_counter()
    while True:
        rc = call()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= max_calls:
            break
    assert rc == 0
    if want_ydot:
        return el / n, n, ydot0
    return el / n, n


def extended_for_oracle(loc):
    """A partition ('owned + halo', shud_up_b200.partition.extract) as one ordinary mesh the CPU oracle can run:
    every halo cell becomes a cell of its own at index Ne+h carrying what the partition knows about it - its statics
    z_surf, z_bottom and effKH parameters (halo_*) and its exchanged state (Ysurf, Ygw) - with no neighbours of its
    own; its remaining parameters are copies of cell 0 (its ydot is garbage and ignored).  ydot of the owned cells,
    reaches and lakes of this mesh is what the partition must produce.  Checker-side only."""
    from shud_up_b200 import partition
    Ne, Nr, Nl = (int(np.asarray(loc[k]).reshape(-1)[0]) for k in ("Ne", "Nr", "Nl"))
    nh = int(np.asarray(loc["halo_z_surf"]).shape[0])
    ext = {}
    for k, v in loc.items():
        v = np.asarray(v)
        if k.startswith(("halo_", "_", "own_")) or k == "y":
            continue
        if k in partition.EDGE_KEYS or k in ("ele_nabr", "ele_lakenabr"):
            a = v.reshape(3, Ne)
            fill = np.zeros((3, nh), dtype=a.dtype) if k in ("ele_nabr", "ele_lakenabr") else np.repeat(a[:, :1], nh, axis=1)
            ext[k] = np.ascontiguousarray(np.concatenate([a, fill], axis=1)).ravel()
        elif (k.startswith("ele_") or k in partition.CELL_DYN) and v.ndim == 1 and v.shape[0] == Ne:
            fill = np.zeros(nh, dtype=v.dtype) if k in ("ele_iLake", "ele_iBC", "ele_iSS") else np.repeat(v[:1], nh)
            ext[k] = np.concatenate([v, fill])
        else:
            ext[k] = v
    for k in partition.HALO_KEYS:
        ext["ele_" + k][Ne:] = np.asarray(loc["halo_" + k])
    hs = np.asarray(loc["halo_state_expected"]).reshape(nh, 2)
    y = np.asarray(loc["y"])
    aq = ext["ele_AquiferDepth"][Ne:]
    ext["y"] = np.concatenate([y[:Ne], hs[:, 0], y[Ne:2 * Ne], 0.1 * np.maximum(aq - hs[:, 1], 0.02), y[2 * Ne:3 * Ne], hs[:, 1],
                               y[3 * Ne:]])
    ext["Ne"] = np.array([Ne + nh], dtype=np.int32)
    if "riv_toLake" in ext and np.any(np.asarray(ext["riv_toLake"]) >= Nl):
        # reaches flowing into a lake held by another partition (partition.extract_cut marks them with index Nl): the
        # oracle gets one dummy lake to pour them into (its ydot is ignored)
        ptr = np.asarray(ext["lake_bathy_ptr"]).astype(np.int32)
        ext["lake_zmin"] = np.concatenate([np.asarray(ext["lake_zmin"], dtype=np.float64), [0.0]])
        ext["lake_NumEleLake"] = np.concatenate([np.asarray(ext["lake_NumEleLake"]).astype(np.int32), [1]]).astype(np.int32)
        ext["lake_bathy_yi"] = np.concatenate([np.asarray(ext["lake_bathy_yi"], dtype=np.float64), [0.0, 1.0, 2.0]])
        ext["lake_bathy_ai"] = np.concatenate([np.asarray(ext["lake_bathy_ai"], dtype=np.float64), [1.0, 1.0, 1.0]])
        ext["lake_bathy_ptr"] = np.concatenate([ptr, [ptr[-1] + 3]]).astype(np.int32)
        ext["y"] = np.concatenate([ext["y"], [1.0]])
        ext["Nl"] = np.array([Nl + 1], dtype=np.int32)
        ext["lakeon"] = np.array([1], dtype=np.int32)
    return ext, Ne, nh


def oracle_parity(mesh, ydot_gpu, nthreads):
    """the GPU ydot of THIS bench mesh (first f() after set_forcing + prime, reference order) against the CPU oracle
    on the same inputs, at the tolerance of the parity tests: |d| <= 1e-12 max(|ydot_ref|, sum |terms|)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import parity
    if "halo_z_surf" in mesh and int(np.asarray(mesh["halo_z_surf"]).shape[0]) > 0:
        ext, Ne, nh = extended_for_oracle(mesh)
    else:
        ext, Ne, nh = mesh, int(mesh["Ne"][0]), 0
    satn = oracle_lib.oracle_prime(ext, ext["y"])
    o = oracle_lib.oracle_rhs(ext, u_satn=satn, qEleE_IC=ext["qEleE_IC_in"], nthreads=nthreads)
    assert o["err"] == 0
    sc = parity.ydot_scale(ext, o)
    NE = Ne + nh
    keep = np.r_[0:Ne, NE:NE + Ne, 2 * NE:2 * NE + Ne, 3 * NE:o["ydot"].size]  # owned cells x3, reaches, lakes
    ref, sc = o["ydot"][keep], sc[keep]
    bad = parity.mismatches(ydot_gpu, ref, sc)
    rel = np.abs(ydot_gpu - ref) / np.maximum(np.maximum(np.abs(ref), sc), 1e-300)
    return {"n": int(ref.size), "n_bad": int(bad.size), "max_rel": float(rel.max()), "rtol": parity.RTOL,
            "max_abs": float(np.abs(ydot_gpu - ref).max()),
            "against": "oracle/shud_oracle.c (bit-exact to the reference f() on the golden cases), same mesh, forcing, "
                       "carried state and y; scale = max(|ydot_ref|, sum |terms| of the balance equation)"}


def reference_arm(a):
    """the reference's own algorithm for this path on the host cores: the serial f() physics with `omp parallel for`
    on its cell / segment / reach loops (oracle port; the as-shipped OpenMP build drops ET and lakes - SURVEY.md 2.1 -
    and needs the basin text inputs, which do not exist on this box).  Rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    steps, warm = a.steps, max(a.warmup, 3)
    ncpu = os.cpu_count() or 1
    if rank != 0:
        return
    mesh = stripe_mesh(1, 0)
    Ne = int(mesh["Ne"][0])
    # W untimed calls, then exactly K timed calls (one call = one step = one f() over the whole 1M-cell mesh)
    per, n = cpu_oracle_time(mesh, ncpu, budget_s=1e9, max_calls=steps, warm_calls=warm)
    per1, _n1 = cpu_oracle_time(mesh, 1, budget_s=3.0, max_calls=3)  # the reference's serial order, one core
    v = Ne / per
    # where the reference tree and its compiled builds exist (the build container), time the reference's own serial
    # and as-shipped OpenMP f() on the three shipped basins beside the port (builds a / b)
    ref_builds = None
    try:
        from tools import time_reference
        ref_builds = time_reference.run(200)
    except Exception:
        ref_builds = None
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": n,
                      "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": config_for(a.gpus),
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncpu, "kind": "port",
                                       "sample": f"{n} f() calls on the full 1M-cell mesh, oracle/shud_oracle.c with OpenMP",
                                       "serial_value": Ne / per1},
                      "reference_builds": ref_builds,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return



def gpu_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warm = a.steps, max(a.warmup, 3)
    ncpu = os.cpu_count() or 1
    import torch
    import torch.distributed as dist
    from shud_up_b200.api import ShudRHS
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # never overrides the caller's setting (the driver reads NCCL's rank lines)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    mesh = stripe_mesh(world, rank)
    Ne, Nr, Ns, Nl = (int(mesh[k][0]) for k in ("Ne", "Nr", "Ns", "Nl"))
    rhs = ShudRHS(mesh, device=local_rank)
    rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
    rhs.prime(mesh["y"])
    st = rhs.torch_stream()
    dev = torch.device(f"cuda:{local_rank}")
    with torch.cuda.stream(st):
        y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).to(dev)
        y = torch.empty_like(y_ref)
        ydot = torch.empty_like(y_ref)
        rhs.to_device_order(y_ref, y)
    st.synchronize()
    hx = None
    if world > 1:
        from shud_up_b200 import partition
        with torch.cuda.stream(st):
            hx = partition.HaloExchange(mesh, dist, dev, cell_perm=rhs.perm()[0], pack_fn=rhs.pack_halo)
            rhs.set_halo_state(hx.halo_state)
        st.synchronize()

    def step():
        # one f().  N>1: pack + post the halo exchange of the boundary cells' (Ysurf, Ygw) over NCCL, run the part of
        # the RHS that needs no halo data (effKH + every interior tile) while it is in flight, then the rest
        if hx is not None:
            with torch.cuda.stream(st):
                hx.start(y)
                rhs.f_interior_dev(0.0, y, ydot)
                rhs.f_boundary_dev(0.0, y, ydot, halo_stream=hx.finish())
        else:
            rhs.f_dev(0.0, y, ydot)

    # ---------------- parity of THIS mesh: the first f() after set_forcing + prime against the CPU oracle -----------
    step()
    with torch.cuda.stream(st):
        rhs.from_device_order(ydot, y_ref)
        ydot_first = y_ref.cpu().numpy().copy()
        y_ref.copy_(torch.from_numpy(np.ascontiguousarray(mesh["y"])))
    st.synchronize()
    par = oracle_parity(mesh, ydot_first, max(1, ncpu // max(world, 1)))
    del ydot_first
    if world > 1:
        pt = torch.tensor([par["n_bad"], par["n"]], dtype=torch.float64, device=dev)
        pm = torch.tensor([par["max_rel"]], dtype=torch.float64, device=dev)
        dist.all_reduce(pt); dist.all_reduce(pm, op=dist.ReduceOp.MAX)
        par.update(n_bad=int(pt[0]), n=int(pt[1]), max_rel=float(pm[0]), ranks=world)
    assert par["n_bad"] == 0, par

    eager_step, step_mode, g, native, p2p = step, "eager", None, False, False
    if hx is not None and os.environ.get("SHUD_BENCH_GRAPH", "1") != "0":
        # N>1: the step is 7 short launches + one NCCL call from Python; capture it (collective included) into one
        # CUDA graph so that the host does one launch per f(), as the single-GPU path does inside the library
        for _ in range(3):
            eager_step()
        barrier_ = lambda: (dist.barrier(), torch.cuda.synchronize())
        barrier_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
            eager_step()
        barrier_()
        def graph_step():
            with torch.cuda.stream(st):
                g.replay()
        # keep whichever launch mode is faster on this box (max over ranks, so every rank decides alike): with 8
        # ranks the replayed NCCL node has been measured slower than the eager collective, with 2-4 ranks faster
        def probe(fn, n=max(steps, 40)):
            # same conditions as the timed region below: as many back-to-back steps, clock sampler running (with 8
            # ranks a short burst of graph replays has been measured 3x faster than a long one)
            for _ in range(5):
                fn()
            barrier_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(local_rank):
                e0.record(st)
                for _ in range(n):
                    fn()
                e1.record(st)
                barrier_()
            t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        t_eager, t_graph = probe(eager_step), probe(graph_step)
        eager_step(); torch.cuda.synchronize(); yd_torch = ydot.clone()   # the state is stationary after warm-up
        # The product path - and the one that is timed: the C library drives the exchange itself, one C call per f()
        # (shud_b200_rhs_exchange_dev), replayed as one CUDA graph.  Two transports: grouped ncclSend / ncclRecv per
        # neighbour over the library's own communicator, and - when the neighbours' buffers can be mapped (CUDA IPC over
        # NVLink) - peer stores + flags with no collective call at all.  Both must give the bits of the torch exchange.
        rhs.comm_init(dist, dev)
        rhs.exchange_plan(*hx.native_plan())
        def native_step():
            rhs.f_exchange_dev(0.0, y, ydot)
        t_nccl = probe(native_step)
        native_step(); torch.cuda.synchronize()
        assert torch.equal(ydot, yd_torch), "library-driven NCCL exchange differs from the torch.distributed one"
        t_p2p, p2p = float("inf"), False
        if os.environ.get("SHUD_P2P", "1") != "0":
            p2p = rhs.p2p_connect(dist, dev)
            if p2p:
                t_p2p = probe(native_step)
                native_step(); torch.cuda.synchronize()
                assert torch.equal(ydot, yd_torch), "peer-to-peer exchange differs from the torch.distributed one"
        del yd_torch
        native = True
        step = native_step
        step_mode = (f"one C call per f() (shud_b200_rhs_exchange_dev) replayed as a CUDA graph; halo transport: "
                     f"{'peer stores + flags over NVLink (CUDA IPC)' if p2p else 'NCCL send/recv issued by the library'}; "
                     f"probes: library p2p {t_p2p * 1e3:.0f} us, library NCCL {t_nccl * 1e3:.0f} us, torch graph "
                     f"{t_graph * 1e3:.0f} us, torch eager {t_eager * 1e3:.0f} us")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    for _ in range(warm):
        step()
    # the parity check above kept the GPU idle for seconds while the CPU oracle ran: W steps (< 1 ms) do not bring the
    # clocks back from their idle state, so the untimed warm-up goes on for half a second of back-to-back steps (the
    # count is fixed, so every rank issues the same number of exchanges)
    for _ in range(4000):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record(st)
        for _ in range(steps):
            step()
        ev1.record(st)
        barrier()
        ms_dev = ev0.elapsed_time(ev1)
        # keep the GPU under the same load long enough for a few 100-ms clock samples; the count depends on
        # `steps` only, so every rank issues the same number of collectives
        for _ in range(max(0, 6000 - steps)):
            step()
        torch.cuda.synchronize()
    code, where = rhs.check()
    assert code == 0, (code, where)
    clocks = clk.summary()

    # ---------------- per-kernel timing (roofline of the dominant kernel) ----------------
    nst = rhs.launches_per_rhs
    kt = []
    for s in range(nst):
        for _ in range(3):
            rhs.f_stage_dev(s, y, ydot)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            rhs.f_stage_dev(s, y, ydot)
        e1.record(st)
        st.synchronize()
        kt.append(e0.elapsed_time(e1) / steps)

    # ---------------- end to end through the CVRhsFn-shaped entry point (host vectors) ----------------
    yh = torch.from_numpy(np.ascontiguousarray(mesh["y"])).pin_memory()
    ydh = torch.empty_like(yh).pin_memory()
    def e2e_step():
        if hx is None:
            rhs.f(0.0, yh, ydh)  # H2D y, permute, 3 kernels, permute, D2H ydot, sync, error word
        else:                    # same sequence with the halo exchange between the upload and the kernels
            with torch.cuda.stream(st):
                y_ref.copy_(yh, non_blocking=True)
                rhs.to_device_order(y_ref, y)
                if native:
                    rhs.f_exchange_dev(0.0, y, ydot)
                else:
                    hx.start(y)
                    rhs.f_interior_dev(0.0, y, ydot)
                    rhs.f_boundary_dev(0.0, y, ydot, halo_stream=hx.finish())
                rhs.from_device_order(ydot, y_ref)
                ydh.copy_(y_ref, non_blocking=True)
            st.synchronize()
            assert rhs.check()[0] == 0
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(steps, 50))
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    s_e2e = (time.perf_counter() - t0) / n_e2e
    if hx is not None:  # restore the device-order state for anything that follows
        with torch.cuda.stream(st):
            y_ref.copy_(yh); rhs.to_device_order(y_ref, y)
        st.synchronize()

    # ---------------- land-surface step on the device (updateforcing + ET), against the upload it replaces --------
    land_rec = None
    if world == 1:
        from shud_up_b200 import abi as _abi
        rng = np.random.default_rng(20240611 + 7)
        tilt = rng.normal(0, 0.15, (2, Ne)); nz_ = 1 / np.sqrt(1 + (tilt ** 2).sum(0))
        lsnap = {"land_nforc": [1], "land_nlc": [12], "land_nmf": [1], "land_iForc": np.ones(Ne, np.int32),
                 "land_iLC": rng.integers(1, 13, Ne).astype(np.int32), "land_iMF": np.ones(Ne, np.int32),
                 "land_Albedo": rng.uniform(0.1, 0.3, Ne), "land_FixPressure": rng.uniform(85, 95, Ne),
                 "land_windH": np.full(Ne, 10.0), "land_nx": tilt[0] * nz_, "land_ny": tilt[1] * nz_, "land_nz": nz_,
                 "land_forc_z": [-9999.0], "land_gc": [1, 0, 1, 1, 1, 1], "land_cs": [0, 1, 0, 5.0, 0.05, 1]}
        Lnd, _k1 = _abi.make_land(lsnap)
        rhs.land_create(Lnd)
        rhs.land_set_state(np.zeros(Ne), np.zeros(Ne))
        S = _abi.ShudLandStep()
        _arr = {"forc": np.array([12.0, 1.5, 0.85, 2.0, 150.0]), "lai": rng.uniform(0.3, 5.0, 12), "mf": np.array([0.0013]),
                "tsr_sx": np.array([-0.6, -0.5]), "tsr_sy": np.array([-0.5, -0.4]), "tsr_sz": np.array([0.62, 0.77]),
                "tsr_wdt": np.array([18.6, 23.1])}
        for k_, v_ in _arr.items():
            setattr(S, k_, v_.ctypes.data_as(_abi._PD))
        S.tsr_n, S.tsr_den, S.dt_min = 2, 41.7, 60.0
        for _ in range(3):
            rhs.land_step(S)
        st.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            rhs.land_step(S)
        st.synchronize()
        ms_land = (time.perf_counter() - t0) / 20 * 1e3
        assert rhs.check()[0] == 0
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a_, b_ in ev:  # device time of one step (k_land_tables + k_land), CUDA events on the context stream
            a_.record(st); rhs.land_step(S); b_.record(st)
        st.synchronize()
        ms_land_dev = float(np.median([a_.elapsed_time(b_) for a_, b_ in ev]))
        t0 = time.perf_counter()
        for _ in range(5):
            rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
        st.synchronize()
        ms_up = (time.perf_counter() - t0) / 5 * 1e3
        rhs.prime(mesh["y"])
        peak_l, _src = measured_peak_gbs()
        land_rec = {"ms_per_step": ms_land, "device_ms": ms_land_dev, "bytes_per_cell": 208,
                    "achieved_gbs": 208 * Ne / (ms_land_dev * 1e-3) / 1e9,
                    "roofline_frac": 208 * Ne / (ms_land_dev * 1e-3) / 1e9 / peak_l,
                    "replaces_upload_ms": ms_up,
                    "what": "shud_b200_land_step: tReadForcing + ET per cell on the device (terrain radiation on); ms_per_step = "
                            "host call to completion, device_ms = k_land_tables + k_land by CUDA events around each call (calls back to back), 208 B/cell "
                            "algorithmic (96 read, 112 written); replaces_upload_ms = shud_b200_set_forcing of the 7 per-cell "
                            "arrays the reference's host loop would hand over each ET step"}

    # ---------------- RHS + SPGMR together (configs[3]): BDF / Newton / SPGMR steps of the library's integrator ------
    # The CVODE-shaped integrator (include/shud_cvode.h) runs in C on SHUD B200 N_Vectors: the time loop, the Newton
    # iteration, SPGMR and every vector operation are library code; Python only starts each step.  The device-fused
    # Newton-Krylov pieces (shud_b200_cv_fused_create) at every N; N > 1: distributed vectors - every reduction is a
    # local kernel + one ncclAllReduce of the scalar on the device inside the library - and f() = halo exchange + RHS
    # (shud_b200_f_exchange).  SHUD_NK_FUSED=0: the generic operations table instead.
    nk = None
    if world == 1 or native:
        import ctypes as C
        from shud_up_b200 import cvode as _cv
        from shud_up_b200.api import lib as _lib
        L = _cv.bind(_lib())
        L.N_VNew_ShudB200.restype = C.c_void_p
        L.N_VNew_ShudB200.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.N_VCopyToDevice_ShudB200.argtypes = [C.c_void_p]
        L.N_VSetDistributed_ShudB200.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        ws = C.c_void_p()
        assert L.shud_nv_ws_create(local_rank, C.c_void_p(rhs.stream_ptr), C.byref(ws)) == 0
        yv = C.c_void_p(L.N_VNew_ShudB200(rhs.NY, ws, rhs._h, None))
        np.ctypeslib.as_array(L.N_VGetArrayPointer(yv), shape=(rhs.NY,))[:] = mesh["y"]
        assert L.N_VCopyToDevice_ShudB200(yv) == 0
        n_glob = rhs.NY
        if world > 1:
            ng = torch.tensor([float(rhs.NY)], dtype=torch.float64, device=dev)
            dist.all_reduce(ng)
            n_glob = int(ng[0])
            L.N_VSetDistributed_ShudB200(yv, n_glob, C.c_void_p(_cv.fn_address(L, "shud_b200_nv_allreduce")), rhs._h)
            _nr, _rk, _bx = C.c_int(0), C.c_int(0), (C.c_void_p * 16)()
            L.shud_b200_p2p_mailboxes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
            L.shud_b200_p2p_mailboxes(rhs._h, C.byref(_nr), C.byref(_rk), _bx)
            ar_mode = ("inside the reduction kernels: partials stored into every rank's mailbox over NVLink, combined in rank order"
                       if _nr.value > 1 and os.environ.get("SHUD_P2P_AR", "1") != "0" else "ncclAllReduce of the scalar on the stream")
        rhs.prime(mesh["y"])
        cvi = _cv.CVode(L, _cv.fn_address(L, "shud_b200_f_exchange" if world > 1 else "shud_b200_f"), rhs._h.value, 0.0, yv)
        cvi.configure(rtol=1e-4, atol=1e-4, init_step=1e-3, max_step=10.0)
        fz = None
        if os.environ.get("SHUD_NK_FUSED", "1") != "0":
            fz = _cv.Fused()
            L.shud_b200_cv_fused_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_cv.Fused)]
            L.shud_b200_cv_fused_destroy.argtypes = [C.POINTER(_cv.Fused)]
            assert L.shud_b200_cv_fused_create(rhs._h, ws, 5, C.byref(fz)) == 0
            cvi.set_fused(fz)
        for _ in range(3):
            cvi.solve(1e9, yv, itask=_cv.CV_ONE_STEP)
        barrier()
        s0 = cvi.stats()
        t_sim0 = cvi.t
        t0 = time.perf_counter()
        for _ in range(20):
            cvi.solve(1e9, yv, itask=_cv.CV_ONE_STEP)
        barrier()
        w_nk = time.perf_counter() - t0
        s1 = cvi.stats()
        nrhs = (s1["nfe"] + s1["nfeLS"]) - (s0["nfe"] + s0["nfeLS"])
        sim_min = cvi.t - t_sim0
        tot_cells = Ne
        if world > 1:
            tc = torch.tensor([float(Ne)], dtype=torch.float64, device=dev)
            dist.all_reduce(tc)
            tot_cells = int(tc[0])
        nk = {"bdf_steps": s1["nst"] - s0["nst"], "rhs_calls": nrhs, "newton_iters": s1["nni"] - s0["nni"],
              "krylov_iters": s1["nli"] - s0["nli"], "order": s1["qlast"], "wall_ms": w_nk * 1e3,
              "cell_updates_per_s": nrhs * tot_cells / w_nk,
              # BASELINE.json's second metric on this mesh: simulated time advanced by these steps (start-up phase of the
              # integration: the step size is still growing from init_step) per wall second
              "sim_minutes": sim_min, "sim_days_per_wall_s": sim_min / 1440.0 / w_nk,
              "ms_per_rhs_call_incl_vector_ops": w_nk * 1e3 / max(nrhs, 1),
              "what": ("CVODE-shaped integrator of the library (BDF 1-5, Newton, SPGMR(5), difference-quotient Jv; "
                       "csrc/shud_cvode.cpp) on SHUD B200 N_Vectors, "
                       + ("device-fused Newton-Krylov pieces (one-pass predictor, shud_spgmr_newton_step)" if fz is not None
                          else "generic operations table")
                       + ("" if world == 1 else
                          f"; {world} partitions: distributed vectors, allreduce of every reduction {ar_mode}, "
                          "f() = halo exchange + RHS")),
              "fused": fz is not None}
        cvi.close()
        if fz is not None:
            L.shud_b200_cv_fused_destroy(C.byref(fz))
        L.N_VDestroy(yv)
        L.shud_nv_ws_destroy(ws)
        with torch.cuda.stream(st):
            rhs.to_device_order(y_ref, y)
        st.synchronize()

    # ---------------- strong-scaling anchor: the 8 x 1M cells of the N=8 weak-scaling run in ONE context -------------
    # (N=1 only, driver-run and clock-sampled like the headline: SCALE's N=8 time against this one is the strong ratio)
    anchor = None
    if world == 1 and os.environ.get("SHUD_BENCH_ANCHOR", "1") != "0":
        from shud_up_b200 import synth as _synth
        try:
            t_a0 = time.perf_counter()
            m8 = _synth.make(**_synth.named("8M"))
            r8 = ShudRHS(m8, device=local_rank)
            r8.set_forcing(m8, qEleE_IC=m8["qEleE_IC_in"])
            r8.prime(m8["y"])
            s8 = r8.torch_stream()
            with torch.cuda.stream(s8):
                y8r = torch.from_numpy(np.ascontiguousarray(m8["y"])).to(dev)
                y8, yd8 = torch.empty_like(y8r), torch.empty_like(y8r)
                r8.to_device_order(y8r, y8)
            s8.synchronize()
            for _ in range(5):
                r8.f_dev(0.0, y8, yd8)
            s8.synchronize()
            n8 = 60
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(local_rank) as clk8:
                a0.record(s8)
                for _ in range(n8):
                    r8.f_dev(0.0, y8, yd8)
                a1.record(s8)
                s8.synchronize()
                for _ in range(1500):  # long enough for a few clock samples
                    r8.f_dev(0.0, y8, yd8)
                s8.synchronize()
            assert r8.check()[0] == 0
            ms8 = a0.elapsed_time(a1) / n8
            ne8 = int(m8["Ne"][0])
            anchor = {"workload": "synthetic-8M in one context: 8,000,000 cells / 400,000 reaches / 1,200,000 segments",
                      "ms_per_step": ms8, "value": ne8 / (ms8 * 1e-3), "unit": UNIT, "steps": n8,
                      "rhs_frac": (B_CELL * ne8 + B_RIV * r8.Nr + B_SEG * r8.Ns) / (ms8 * 1e-3) / 1e9 / measured_peak_gbs()[0],
                      "clocks": clk8.summary(), "setup_s": time.perf_counter() - t_a0 - ms8 * 1e-3 * (n8 + 1505)}
            r8.close()
            del m8, y8r, y8, yd8
        except Exception as exc:  # the anchor must never cost the headline
            anchor = {"unavailable": repr(exc)[:200]}

    # ---------------- reduce over ranks: max time ----------------
    tt = torch.tensor([ms_dev, s_e2e * 1e3], dtype=torch.float64, device=dev)
    cells = torch.tensor([float(Ne)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    ms_dev_max, ms_e2e_max = float(tt[0]), float(tt[1])
    total_cells = float(cells[0])
    per_rank = [ms_dev / steps]
    if world > 1:  # the per-rank device times (the reported one is their max) and tile split, for the record
        allt = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
        n_int, n_bnd = rhs.tile_counts()
        dist.all_gather(allt, torch.tensor([ms_dev / steps, n_int, n_bnd], dtype=torch.float64, device=dev))
        per_rank = [[round(float(a[0]), 5), int(a[1]), int(a[2])] for a in allt]
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_step = ms_dev_max / steps
        value = total_cells * steps / (ms_dev_max * 1e-3)
        b_rhs = B_CELL * Ne + B_RIV * Nr + B_SEG * Ns
        dom = int(np.argmax(kt))
        names = ["k_effkh", "k_fused", "k_river_lake"]
        # bytes of the dominant (cell) kernel: everything per cell and per segment except what the effKH
        # pre-pass alone touches (its 4 parameters + its effKH store: 40 B/cell), DESIGN.md section 4
        b_dom = {0: 60 * Ne, 1: (B_CELL - 40) * Ne + B_SEG * Ns, 2: B_RIV * Nr + 16 * Ns}[dom]
        ach = b_dom / (kt[dom] * 1e-3) / 1e9
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": config_for(world),
               "multi_gpu_detail": (f"halo exchange of {hx.bytes_per_exchange} B per rank and f() "
                                    + ("inside the three launches of f() (pre-pass stores into the neighbours' buffers, halo "
                                       "tiles acquire the flags)" if p2p else "overlapped with the interior tiles")
                                    + f"; step launched as: {step_mode}; per rank [ms/step, interior tiles, boundary "
                                    f"tiles]: {per_rank}") if world > 1 else None,
               "parity": par,
               # peer-to-peer: the same three kernels as a single domain; NCCL path: + pack, the split cell kernel, halo pre-pass
               "gpu_launches": (nst + ((0 if p2p else 2) if world > 1 else 0)) * steps,
               "clocks": clocks,
               "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": ncu_traffic(names[dom]), "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": b_dom, "kernel_ms": kt[dom],
                            "all_kernels_ms": dict(zip(names, kt)),
                            "rhs_bytes": b_rhs, "rhs_achieved_gbs": b_rhs / (ms_step * 1e-3) / 1e9,
                            "rhs_frac": b_rhs / (ms_step * 1e-3) / 1e9 / peak},
               "e2e": {"value": total_cells / (ms_e2e_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * rhs.NY * world,
                       "d2h_bytes_per_step": 8 * rhs.NY * world, "ms_per_step": ms_e2e_max}}
        if anchor is not None:
            out["strong_scaling_anchor"] = anchor
        if nk is not None:
            out["newton_krylov"] = nk
        if land_rec is not None:
            out["land_surface_step"] = land_rec
        if world == 1:
            per, n = cpu_oracle_time(mesh, ncpu, budget_s=a.cpu_budget)
            per1, n1 = cpu_oracle_time(mesh, 1, budget_s=min(4.0, a.cpu_budget), max_calls=5)
            out["cpu_baseline"] = {"value": Ne / per, "unit": UNIT, "cores": ncpu, "kind": "port",
                                   "sample": f"{n} f() calls on the full 1M-cell mesh, oracle/shud_oracle.c with OpenMP on {ncpu} threads",
                                   "ms_per_step": per * 1e3,
                                   # the reference's serial order exactly (SURVEY.md 8(d) build 1), one core
                                   "serial_value": Ne / per1, "serial_ms_per_step": per1 * 1e3, "serial_calls": n1}
        print(json.dumps(out), flush=True)
    return rhs, g, hx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    import gc
    import torch
    import torch.distributed as dist
    rhs, g, hx = gpu_arm(a)
    # Orderly teardown, then leave through the interpreter's normal exit path (no os._exit).  Everything gpu_arm
    # allocated on the context's stream died with its frame; what is left goes in dependency order: graph ->
    # exchange buffers -> cached blocks of that stream -> the context (its streams, NCCL communicator) -> process group.
    torch.cuda.synchronize()
    if g is not None:
        g.reset()
    del g, hx
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    rhs.close()
    del rhs
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()


if __name__ == "__main__":
    main()
