"""Partitioned RHS on the GPU: two partition contexts (both on cuda:0, the exchange done in-process through
the same pack kernel and buffers the NCCL path uses) reproduce the single-domain CUDA RHS bit for bit on the
owned cells - ydot needs no cross-partition sum (owner-computes)."""
import numpy as np
import pytest

import oracle_lib
import parity
from shud_up_b200 import partition, synth

pytestmark = pytest.mark.gpu
NX, NY, NT, RPT = 60, 40, 4, 60


def _run(mesh, halo_state=None, hx=None):
    import torch
    from shud_up_b200.api import ShudRHS
    rhs = ShudRHS(mesh)
    rhs.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
    rhs.prime(mesh["y"])
    st = rhs.torch_stream()
    with torch.cuda.stream(st):
        y_ref = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda()
        y, yd, yd_ref = torch.empty_like(y_ref), torch.empty_like(y_ref), torch.empty_like(y_ref)
        rhs.to_device_order(y_ref, y)
    return rhs, st, y, yd, yd_ref


def test_two_partitions_equal_single_domain_bitwise():
    import torch
    whole = synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT)
    Ne = int(whole["Ne"][0])
    rhs, st, y, yd, yd_ref = _run(whole)
    with torch.cuda.stream(st):
        rhs.f_dev(0.0, y, yd)
        rhs.from_device_order(yd, yd_ref)
    st.synchronize()
    assert rhs.check()[0] == 0
    ref = yd_ref.cpu().numpy()
    # and the single-domain CUDA result agrees with the oracle
    w2 = dict(whole); w2["ele_u_satn"] = oracle_lib.oracle_prime(whole, whole["y"])
    o = oracle_lib.oracle_rhs(w2)
    assert parity.mismatches(ref, o["ydot"], parity.ydot_scale(w2, o)).size == 0
    gid_w = whole["ele_gid"]
    pos = np.empty(gid_w.max() + 1, dtype=np.int64); pos[gid_w] = np.arange(Ne)
    locs = [synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT, rows=(r0, r0 + 20), stripe_rows=20) for r0 in (0, 20)]
    ctxs = [_run(l) for l in locs]
    all_halo = [l["halo_gid"] for l in locs]
    # exchange plan exactly as HaloExchange builds it, the all_to_all replaced by a device copy
    plans = [partition.exchange_plan(l["own_gid"], l["halo_gid"], all_halo) for l in locs]
    for p, (loc, (r, s, yy, ydd, ydr)) in enumerate(zip(locs, ctxs)):
        q = 1 - p
        rq, sq, yq = ctxs[q][0], ctxs[q][1], ctxs[q][2]
        ids_q = plans[q][0][p]                                   # reference-local ids rank q sends to p
        inv = np.empty(rq.Ne, dtype=np.int64); inv[rq.perm()[0]] = np.arange(rq.Ne)
        idx = torch.from_numpy(inv[ids_q].astype(np.int32)).cuda()
        buf = torch.zeros(2 * idx.numel(), dtype=torch.float64, device="cuda")
        with torch.cuda.stream(sq):
            rq.pack_halo(yq, idx, buf)
        sq.synchronize()
        assert np.array_equal(buf.cpu().numpy(), loc["halo_state_expected"])
        r.set_halo_state(buf)
        with torch.cuda.stream(s):
            r.f_dev(0.0, yy, ydd)
            r.from_device_order(ydd, ydr)
        s.synchronize()
        assert r.check()[0] == 0
        got = ydr.cpu().numpy()
        ne = r.Ne
        sel = pos[loc["own_gid"]]
        for b in range(3):
            assert np.array_equal(got[b * ne:(b + 1) * ne], ref[b * Ne + sel]), (p, b)
        # the interior / boundary split (what overlaps the NCCL exchange) gives the same bits
        n_int, n_bnd = r.tile_counts()
        assert n_int > 0 and n_bnd > 0 and n_int + n_bnd == (ne + 127) // 128
        r.prime(loc["y"]); r.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
        yd2 = torch.full_like(ydd, float("nan"))
        with torch.cuda.stream(s):
            r.f_interior_dev(0.0, yy, yd2)
            r.f_boundary_dev(0.0, yy, yd2)
            r.from_device_order(yd2, ydr)
        s.synchronize()
        assert r.check()[0] == 0
        assert np.array_equal(ydr.cpu().numpy(), got)
        # ... and with the halo-dependent tiles on a second stream (where the exchange completes)
        side = torch.cuda.Stream(priority=-1)
        r.prime(loc["y"]); r.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
        yd3 = torch.full_like(ydd, float("nan"))
        with torch.cuda.stream(s):
            r.f_interior_dev(0.0, yy, yd3)
            r.f_boundary_dev(0.0, yy, yd3, halo_stream=side)
            r.from_device_order(yd3, ydr)
        s.synchronize()
        assert r.check()[0] == 0
        assert np.array_equal(ydr.cpu().numpy(), got)
    # ... and with the peer-to-peer exchange of the library: each partition's pack kernel stores its boundary cells
    # straight into the other's halo buffer and releases a flag, the boundary tiles start behind the flag wait
    # (contexts of one process are connected by pointer; across processes the same buffers are mapped through CUDA
    # IPC).  Several calls in a row: the two halo buffers alternate, the epochs advance in step.
    for p, (loc, (r, s, yy, ydd, ydr)) in enumerate(zip(locs, ctxs)):
        q = 1 - p
        r.exchange_plan(np.array([q]), np.array([len(plans[p][0][q])]), np.array([len(loc["halo_gid"])]), plans[p][0][q])
    blobs = [ctxs[p][0].p2p_export(p) for p in range(2)]
    for p in range(2):
        assert ctxs[p][0].p2p_connect_blobs(p, blobs)
    torch.cuda.synchronize()
    for it in range(5):
        outs = []
        for p, (loc, (r, s, yy, ydd, ydr)) in enumerate(zip(locs, ctxs)):
            r.prime(loc["y"]); r.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
            outs.append(torch.full_like(ydd, float("nan")))
        torch.cuda.synchronize()
        for p, (loc, (r, s, yy, ydd, ydr)) in enumerate(zip(locs, ctxs)):
            r.f_exchange_dev(0.0, yy, outs[p])      # asynchronous: partition 0 waits on the device for partition 1's flag
        for p, (loc, (r, s, yy, ydd, ydr)) in enumerate(zip(locs, ctxs)):
            with torch.cuda.stream(s):
                r.from_device_order(outs[p], ydr)
            s.synchronize()
            assert r.check()[0] == 0
            got = ydr.cpu().numpy()
            sel = pos[loc["own_gid"]]
            for b in range(3):
                assert np.array_equal(got[b * r.Ne:(b + 1) * r.Ne], ref[b * Ne + sel]), (it, p, b)
    # reaches: both partitions own whole trees; their ydot equals the single-domain one (same tree order)
    nr0 = ctxs[0][0].Nr
    assert np.array_equal(np.r_[ctxs[0][4].cpu().numpy()[3 * ctxs[0][0].Ne:], ctxs[1][4].cpu().numpy()[3 * ctxs[1][0].Ne:]],
                          ref[3 * Ne:])


@pytest.mark.parametrize("basin,case,nparts", [("qhh", "rand4", 2), ("ccw", "rand1", 3), ("heihe", "rand3", 2)])
def test_assigned_partitions_of_real_basins_on_the_gpu(basin, case, nparts):
    """partition.assign -> partition.extract -> one context per partition (all on cuda:0), halo state taken from the
    whole state as the exchange delivers it: the owned cells, reaches and lakes get the bits of the single-domain
    CUDA RHS (qhh: the lake and its banks live in one partition; heihe: reaches without segments follow their tree)"""
    import torch
    whole = oracle_lib.load_case(basin, case)
    Ne, Nr, Nl = int(whole["Ne"][0]), int(whole["Nr"][0]), int(whole["Nl"][0])

    def run(mesh, halo_pairs=None):
        from shud_up_b200.api import ShudRHS
        r = ShudRHS(mesh)
        r.set_forcing(mesh, qEleE_IC=mesh["qEleE_IC_in"])
        r.set_carried(mesh["ele_u_satn"])
        s = r.torch_stream()
        if halo_pairs is not None:
            hb = torch.from_numpy(np.ascontiguousarray(halo_pairs)).cuda()
            r.set_halo_state(hb)
        with torch.cuda.stream(s):
            yr = torch.from_numpy(np.ascontiguousarray(mesh["y"])).cuda()
            yd, ydd, out = torch.empty_like(yr), torch.empty_like(yr), torch.empty_like(yr)
            r.to_device_order(yr, yd)
            r.f_dev(0.0, yd, ydd)
            r.from_device_order(ydd, out)
        s.synchronize()
        assert r.check()[0] == 0
        res = out.cpu().numpy()
        r.close()
        return res

    ref = run(whole)
    y = np.asarray(whole["y"])
    part = partition.assign(whole, nparts)
    for p in range(nparts):
        loc = partition.extract(whole, part == p, part_of_cell=part, keep_full_halo=True)
        ne, nr, nl = int(loc["Ne"][0]), int(loc["Nr"][0]), int(loc["Nl"][0])
        halo = loc["_halo_ref"]
        pairs = np.stack([y[halo], y[2 * Ne + halo]], 1).ravel()
        got = run(loc, pairs)
        own = loc["_own_ref"]
        for b in range(3):
            assert np.array_equal(got[b * ne:(b + 1) * ne], ref[b * Ne + own]), (p, b)
        assert np.array_equal(got[3 * ne:3 * ne + nr], ref[3 * Ne + loc["_riv_ref"]])
        if nl:
            assert np.array_equal(got[3 * ne + nr:], ref[3 * Ne + Nr:])


@pytest.mark.parametrize("basin,case,nparts", [("ccw", "rand1", 4), ("heihe", "rand3", 3), ("heihe", "rand3", 8),
                                               ("qhh", "rand4", 4), ("qhh", "lakes6", 3)])
def test_cut_river_trees_on_the_gpu(basin, case, nparts):
    """Hilbert-range partitions that cut the river network anywhere (partition.assign_cells / extract_cut): every rank's
    context holds own + ghost cells and own + ghost reaches, the peer-to-peer exchange delivers halo pairs, ghost-cell
    triples and ghost-reach stages, and the owned cells, reaches and lakes get the bits of the single-domain CUDA RHS.
    All contexts on cuda:0, connected by pointer; three calls in a row (alternating halo buffers, advancing epochs)."""
    import torch
    mesh = oracle_lib.load_case(basin, case)
    Ne, Nr, Nl = int(mesh["Ne"][0]), int(mesh["Nr"][0]), int(mesh["Nl"][0])
    mesh = dict(mesh); mesh["ele_u_satn"] = oracle_lib.oracle_prime(mesh, mesh["y"])
    rhs, st, y, yd, yd_ref = _run(mesh)
    with torch.cuda.stream(st):
        rhs.f_dev(0.0, y, yd)
        rhs.from_device_order(yd, yd_ref)
    st.synchronize()
    assert rhs.check()[0] == 0
    ref = yd_ref.cpu().numpy()
    part = partition.assign_cells(mesh, nparts)
    closures = [partition._closure_with_lakes(mesh, part, p) for p in range(nparts)]
    ex = [partition.extract_cut(mesh, part, p, closures) for p in range(nparts)]
    ctxs = [_run(loc) for loc, _ in ex]
    assert sum(int(loc["n_ghost_cells"][0]) + int(loc["n_ghost_reaches"][0]) for loc, _ in ex) > 0
    for (loc, plan), c in zip(ex, ctxs):
        c[0].exchange_plan_items(plan)
    blobs = [ctxs[p][0].p2p_export(p) for p in range(nparts)]
    for p in range(nparts):
        assert ctxs[p][0].p2p_connect_blobs(p, blobs)
    torch.cuda.synchronize()
    for it in range(3):
        outs = []
        for (loc, plan), (r, s, yy, ydd, ydr) in zip(ex, ctxs):
            r.prime(loc["y"]); r.set_forcing(loc, qEleE_IC=loc["qEleE_IC_in"])
            outs.append(torch.full_like(ydd, float("nan")))
        torch.cuda.synchronize()
        for p, (r, s, yy, ydd, ydr) in enumerate(ctxs):
            r.f_exchange_dev(0.0, yy, outs[p])
        for p, ((loc, plan), (r, s, yy, ydd, ydr)) in enumerate(zip(ex, ctxs)):
            with torch.cuda.stream(s):
                r.from_device_order(outs[p], ydr)
            s.synchronize()
            assert r.check()[0] == 0
            got = ydr.cpu().numpy()
            nloc, nown, nro, nl = r.Ne, loc["_own_ref"].size, loc["_riv_ref"].size, loc["_lake_ref"].size
            ngc, ngr = int(loc["n_ghost_cells"][0]), int(loc["n_ghost_reaches"][0])
            for b in range(3):
                assert np.array_equal(got[b * nloc:b * nloc + nown], ref[b * Ne + loc["_own_ref"]]), (it, p, b)
                assert np.all(got[b * nloc + nown:(b + 1) * nloc] == 0.0)                       # ghost cells: ydot 0
            assert np.array_equal(got[3 * nloc:3 * nloc + nro], ref[3 * Ne + loc["_riv_ref"]]), (it, p)
            assert np.all(got[3 * nloc + nro:3 * nloc + r.Nr] == 0.0)                           # ghost reaches: ydot 0
            assert np.array_equal(got[3 * nloc + r.Nr:3 * nloc + r.Nr + nl], ref[3 * Ne + Nr + loc["_lake_ref"]]), (it, p)
    for c in ctxs:
        c[0].close()
    rhs.close()
