"""Multi-GPU host logic on the CPU: stripes of the synthetic mesh generated on their own equal the same part
of the whole mesh; the partition + halo exchange reproduces the single-domain RHS bit for bit (checked with
the CPU oracle on 'owned + halo' meshes); the exchange itself runs over torch.distributed gloo, world_size 2."""
import os
import sys

import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import partition, synth

NX, NY, NT, RPT = 40, 40, 4, 40


@pytest.fixture(scope="module")
def whole():
    return synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT)


def _stripes():
    return [synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT, rows=(r0, r0 + 20), stripe_rows=20) for r0 in (0, 20)]


def test_stripe_generated_alone_equals_that_part_of_the_whole_mesh(whole):
    gid_w = whole["ele_gid"]
    pos = np.empty(gid_w.max() + 1, dtype=np.int64); pos[gid_w] = np.arange(gid_w.size)
    tot = 0
    for loc in _stripes():
        sel = pos[loc["own_gid"]]
        tot += sel.size
        for k in ("ele_area", "ele_z_surf", "ele_AquiferDepth", "ele_KsatH", "ele_Beta", "qPotEvap", "t_lai", "qEleE_IC_in"):
            assert np.array_equal(np.sort(loc[k]), np.sort(whole[k][sel])), k
        ne = int(loc["Ne"][0])
        # state of owned cells, matched through the global id
        o = np.argsort(loc["own_gid"]); w = np.argsort(gid_w[sel])
        Ne_w = int(whole["Ne"][0])
        for b in range(3):
            assert np.array_equal(loc["y"][b * ne:(b + 1) * ne][o], whole["y"][b * Ne_w + sel][w])
        assert loc["halo_gid"].size == NX and int(loc["Nr"][0]) == 2 * RPT
    assert tot == gid_w.size


def _extended(mesh, loc):
    """'owned + halo' as one ordinary mesh the CPU oracle can run: halo cells become real cells whose own
    far-side neighbours are cut off (their ydot is garbage and ignored)."""
    Ne = int(mesh["Ne"][0])
    own, halo = loc["_own_ref"], loc["_halo_ref"]
    cells = np.concatenate([own, halo])
    new = np.zeros(Ne, dtype=np.int64); new[cells] = np.arange(1, cells.size + 1)
    ext = {k: v for k, v in loc.items() if not k.startswith(("halo_", "_", "own_"))}
    for k, v in mesh.items():
        v = np.asarray(v)
        if k in partition.EDGE_KEYS:
            ext[k] = np.ascontiguousarray(v.reshape(3, Ne)[:, cells]).ravel()
        elif (k.startswith("ele_") or k in partition.CELL_DYN) and v.ndim == 1 and v.shape[0] == Ne:
            ext[k] = v[cells]
    nab = np.asarray(mesh["ele_nabr"]).reshape(3, Ne)[:, cells]
    ext["ele_nabr"] = np.where(nab > 0, new[np.maximum(nab - 1, 0)], 0).astype(np.int32).ravel()
    # lake banks are never cut (partition.extract refuses): owned cells keep their lake neighbours, halo cells have none
    lk = np.zeros((3, cells.size), dtype=np.int32)
    lk[:, :own.size] = np.asarray(loc["ele_lakenabr"]).reshape(3, own.size)
    ext["ele_lakenabr"] = lk.ravel()
    ext["Ne"] = np.array([cells.size], dtype=np.int32)
    return ext, own.size, cells.size


def test_partitioned_rhs_equals_single_domain_bitwise(whole):
    Ne, Nr = int(whole["Ne"][0]), int(whole["Nr"][0])
    whole = dict(whole)
    whole["ele_u_satn"] = oracle_lib.oracle_prime(whole, whole["y"])
    ref = oracle_lib.oracle_rhs(whole, want_diag=False)["ydot"]
    row = whole["ele_gid"] // (2 * NX)
    part = row // 20
    locs = [partition.extract(whole, part == p, gid=whole["ele_gid"], part_of_cell=part, keep_full_halo=True) for p in (0, 1)]
    # in-process emulation of the exchange: what rank p receives is what the owners hold
    send_ids = [partition.exchange_plan(l["own_gid"], l["halo_gid"], [x["halo_gid"] for x in locs]) for l in locs]
    for p, loc in enumerate(locs):
        ne = int(loc["Ne"][0])
        q = 1 - p
        ids_q, gids_q = send_ids[q][0][p], send_ids[q][1][p]          # what q sends to p, ordered by gid
        yq, neq = locs[q]["y"], int(locs[q]["Ne"][0])
        msg = np.stack([yq[ids_q], yq[2 * neq + ids_q]], 1).ravel()
        pos = partition.recv_positions(loc["halo_gid"], gids_q)
        assert np.array_equal(pos, np.arange(pos.size))               # (owner, gid) order: lands in place
        assert np.array_equal(msg, loc["halo_state_expected"])
        # the RHS of 'owned + halo' with the exchanged halo state == the single-domain RHS on the owned cells
        ext, nown, ntot = _extended(whole, loc)
        nh = ntot - nown
        y = np.concatenate([np.r_[loc["y"][0:ne], msg[0::2]], np.r_[loc["y"][ne:2 * ne], np.zeros(nh)],
                            np.r_[loc["y"][2 * ne:3 * ne], msg[1::2]], loc["y"][3 * ne:]])
        ext["y"] = y
        ext["ele_u_satn"] = np.r_[loc["ele_u_satn"], np.zeros(nh)]
        out = oracle_lib.oracle_rhs(ext, want_diag=False)["ydot"]
        own = loc["_own_ref"]
        for b in range(3):
            assert np.array_equal(out[b * ntot:b * ntot + nown], ref[b * Ne + own]), (p, b)
        nr = int(loc["Nr"][0])
        assert np.array_equal(out[3 * ntot:3 * ntot + nr], ref[3 * Ne + loc["_riv_ref"]])


@pytest.mark.parametrize("basin,case,nparts", [("ccw", "rand1", 2), ("ccw", "rand1", 3), ("heihe", "rand3", 2),
                                               ("qhh", "rand4", 2), ("qhh", "rand4", 3)])
def test_assigned_partitions_of_real_basins_equal_single_domain_bitwise(basin, case, nparts):
    """partition.assign keeps river trees, lakes with their banks and head-BC cells whole, so partition.extract
    accepts its partitions; each partition (owned + halo, halo state as the exchange delivers it) reproduces the
    single-domain RHS bit for bit on its cells, reaches and lakes; every cell, reach and lake has exactly one owner"""
    _check_assigned(oracle_lib.load_case(basin, case), nparts)


@pytest.mark.parametrize("kw,nparts", [(dict(nx=40, ny=30, ntree=3, reaches_per_tree=30, lake_frac=0.02), 3),
                                       (dict(nx=24, ny=24, ntree=1, reaches_per_tree=40, lake_frac=0.05), 5),
                                       (dict(nx=50, ny=20, ntree=5, reaches_per_tree=20), 4)])
def test_assigned_partitions_of_synthetic_meshes_with_lakes(kw, nparts):
    kw = dict(kw)
    whole = synth.make(kw.pop("nx"), kw.pop("ny"), **kw)
    whole["ele_u_satn"] = oracle_lib.oracle_prime(whole, whole["y"])
    _check_assigned(whole, nparts)


def _check_assigned(whole, nparts):
    Ne, Nr, Nl = int(whole["Ne"][0]), int(whole["Nr"][0]), int(whole["Nl"][0])
    ref = oracle_lib.oracle_rhs(whole, want_diag=False)["ydot"]
    part = partition.assign(whole, nparts)
    assert part.shape == (Ne,) and set(np.unique(part)) == set(range(nparts))
    seen_cells, seen_riv, seen_lakes = np.zeros(Ne, int), np.zeros(Nr, int), 0
    y = np.asarray(whole["y"])
    for p in range(nparts):
        loc = partition.extract(whole, part == p, part_of_cell=part, keep_full_halo=True)
        ne, nr, nl = int(loc["Ne"][0]), int(loc["Nr"][0]), int(loc["Nl"][0])
        own, halo = loc["_own_ref"], loc["_halo_ref"]
        seen_cells[own] += 1; seen_riv[loc["_riv_ref"]] += 1; seen_lakes += nl
        ext, nown, ntot = _extended(whole, loc)
        nh = ntot - nown
        ext["y"] = np.concatenate([np.r_[loc["y"][0:ne], y[halo]], np.r_[loc["y"][ne:2 * ne], np.zeros(nh)],
                                   np.r_[loc["y"][2 * ne:3 * ne], y[2 * Ne + halo]], loc["y"][3 * ne:]])
        ext["ele_u_satn"] = np.r_[loc["ele_u_satn"], np.zeros(nh)]
        out = oracle_lib.oracle_rhs(ext, want_diag=False)["ydot"]
        for b in range(3):
            assert np.array_equal(out[b * ntot:b * ntot + nown], ref[b * Ne + own]), (p, b)
        assert np.array_equal(out[3 * ntot:3 * ntot + nr], ref[3 * Ne + loc["_riv_ref"]])
        if nl:
            assert np.array_equal(out[3 * ntot + nr:], ref[3 * Ne + Nr:])
    assert np.all(seen_cells == 1) and np.all(seen_riv == 1) and seen_lakes == Nl


def test_cut_river_is_refused(whole):
    row = whole["ele_gid"] // (2 * NX)
    with pytest.raises(NotImplementedError):
        partition.extract(whole, row < 5)      # row 5 is the main stem of the first tree


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    loc = synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT, rows=(20 * rank, 20 * rank + 20), stripe_rows=20)
    hx = partition.HaloExchange(loc, dist, torch.device("cpu"))
    y = torch.from_numpy(loc["y"].copy())
    got = hx.exchange(y)[:2 * hx.nr].numpy().copy()
    ok = bool(np.array_equal(got, loc["halo_state_expected"])) and hx.in_place and hx.nr == NX
    # the start()/finish() form that overlaps the exchange with the interior tiles gives the same halo state
    hx.start(y)
    hx.finish()
    ok = ok and bool(np.array_equal(hx.halo_state[:2 * hx.nr].numpy(), loc["halo_state_expected"]))
    # plan handed to the library-driven exchange (shud_b200_exchange_plan): one peer, NX cells each way, the cells
    # sent are owned cells whose global ids are exactly the peer's halo ids
    peers, sc, rc, cells = hx.native_plan()
    ok = ok and peers.tolist() == [1 - rank] and sc.tolist() == [NX] and rc.tolist() == [NX] and cells.size == NX
    all_halo = [None, None]
    dist.all_gather_object(all_halo, np.asarray(loc["halo_gid"]))
    ok = ok and bool(np.array_equal(np.asarray(loc["own_gid"])[cells], np.sort(all_halo[1 - rank])))
    # the distributed WRMS norm: local sum of squares + allreduce + global length
    w = torch.full_like(y, 0.5)
    s = torch.tensor([float(((y * w) ** 2).sum())], dtype=torch.float64)
    n = torch.tensor([float(y.numel())], dtype=torch.float64)
    dist.all_reduce(s); dist.all_reduce(n)
    q.put((rank, ok, float(torch.sqrt(s / n))))
    dist.destroy_process_group()


def _gloo_nvec_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from host_model import HostOps
    from shud_up_b200.nvector import DistributedOps
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    xg, yg, wg = rng.standard_normal(1001), rng.standard_normal(1001), rng.uniform(0.5, 2, 1001)
    sl = slice(0, 400) if rank == 0 else slice(400, 1001)      # uneven partition of one global vector
    ops = DistributedOps(HostOps(), dist, torch.device("cpu"), n_global=1001)
    x, y, w = xg[sl].copy(), yg[sl].copy(), wg[sl].copy()
    res = dict(dot=ops.N_VDotProd(x, y), wrms=ops.N_VWrmsNorm(x, w), mx=ops.N_VMaxNorm(x), mn=ops.N_VMin(x),
               l1=ops.N_VL1Norm(x), multi=ops.N_VDotProdMulti(x, [y, w]).tolist())
    ref = dict(dot=float(xg @ yg), wrms=float(np.sqrt(np.sum((xg * wg) ** 2) / 1001)), mx=float(np.abs(xg).max()),
               mn=float(xg.min()), l1=float(np.abs(xg).sum()), multi=[float(xg @ yg), float(xg @ wg)])
    ok = all(np.allclose(res[k], ref[k], rtol=1e-13) for k in ref)
    # the reducing ops the local table would answer with a per-rank value: masks, quotients, tests, Newton
    idg = (rng.uniform(size=1001) > 0.3).astype(np.float64)
    deng = np.where(rng.uniform(size=1001) > 0.2, rng.uniform(0.5, 2, 1001), 0.0)
    cg = rng.integers(-2, 3, 1001).astype(np.float64)
    zg = xg.copy(); zg[17] = 0.0                       # one zero, on rank 0 only
    host = HostOps()
    z1, z2, m1, m2 = np.zeros(1001), np.zeros_like(x), np.zeros(1001), np.zeros_like(x)
    res2 = dict(wm=ops.N_VWrmsNormMask(x, w, idg[sl].copy()), mq=ops.N_VMinQuotient(x, deng[sl].copy()),
                it=ops.N_VInvTest(zg[sl].copy(), z2), cm=ops.N_VConstrMask(cg[sl].copy(), x, m2),
                va=ops.N_VWrmsNormVectorArray([x, y], [w, w]).tolist())
    ref2 = dict(wm=host.N_VWrmsNormMask(xg, wg, idg), mq=host.N_VMinQuotient(xg, deng), it=host.N_VInvTest(zg, z1),
                cm=host.N_VConstrMask(cg, xg, m1), va=[host.N_VWrmsNorm(xg, wg), host.N_VWrmsNorm(yg, wg)])
    ok = ok and all(np.allclose(res2[k], ref2[k], rtol=1e-13) for k in ref2) and np.array_equal(m2, m1[sl])
    yy, ac = y.copy(), np.zeros_like(y)
    dele = ops.NewtonUpdate(x, w, yy, ac)
    ok = ok and np.isclose(dele, ref["wrms"], rtol=1e-13) and np.array_equal(yy, y + x) and np.array_equal(ac, x)
    try:
        ops.N_VSomethingNew
        ok = False                                     # unknown ops must not fall through to the local table
    except AttributeError:
        pass
    # the integrator on the distributed table takes the same steps on every rank and the same steps as one domain
    from shud_up_b200.integrator import BDFKrylov
    lam = np.linspace(0.01, 2.0, 1001)

    def run(o, sl_):
        lm = lam[sl_]
        integ = BDFKrylov(o, lambda: np.zeros(lm.size), lambda t, yv, yd: np.copyto(yd, -lm * yv), lm.size, rtol=1e-6,
                          atol=1e-8, max_step=1.0, init_step=1e-3, n_global=1001)
        integ.init(0.0, np.ones(lm.size))
        hs = []
        while integ.t < 3.0 - 1e-12:
            integ.step(3.0); hs.append(integ.t)
        return hs, integ.hist[0].copy()
    hs_d, y_d = run(ops, sl)
    hs_1, y_1 = run(host, slice(0, 1001))
    ok = ok and len(hs_d) == len(hs_1) and np.allclose(hs_d, hs_1, rtol=1e-9) and np.allclose(y_d, y_1[sl], rtol=1e-8)
    ok = ok and np.allclose(y_1, np.exp(-lam * 3.0), atol=1e-4)
    q.put((rank, ok, res["dot"], hs_d))
    dist.destroy_process_group()


def test_distributed_nvector_reductions_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_nvec_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2]          # every rank holds the same global value
    assert res[0][3] == res[1][3]          # ... and takes the same time steps


def test_halo_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] > 0
