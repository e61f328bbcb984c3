"""Synthetic SHUD domains of a named size (BASELINE.json configs 4/5):
a structured-triangulated rectangle (nx x ny quads split alternately, Ne = 2 nx ny), node spacing 100 m
with +-20 % jitter, 16 soil / 16 geology / 12 land-cover classes drawn from the ranges of the ccw tables,
dendritic river trees following grid lines down-slope (3 river-element segments per reach), optional lake.
Seed 20240611.

Every random field is drawn per grid ROW from its own stream (seed, kind, row), so any horizontal stripe of
the global mesh can be generated on its own and is identical to that part of the whole mesh: `make(...,
rows=(r0, r1))` returns the partition owning quad rows [r0, r1) with its halo cells (multi-GPU runs build
only their own stripe).  The river-tree bands never straddle a stripe boundary that is a multiple of the band
height, so reaches are never cut.

Output: dict of numpy arrays with the snapshot naming of oracle/ref_driver.cpp, i.e. the static arrays
Model_Data::initialize() would hold for such a mesh (geometry per _Element::applyGeometry / InitElement /
applyNabor, src/classes/Element.cpp:62-270; river hand-over per _River::updateFrDownstream,
src/classes/River.cpp:74-90), plus one forcing step and a state vector with every branch populated.
Host-side set-up only: no RHS arithmetic here.
"""
import numpy as np

SEED = 20240611
MINRIVSLOPE = 4e-4  # src/Model/Macros.hpp:47


def named(size):
    """the two named benchmark domains"""
    if size in ("1M", "synthetic-1M"):
        return dict(nx=1000, ny=500, ntree=50, reaches_per_tree=1000)    # Ne 1 000 000, Nr 50 000, Ns 150 000
    if size in ("8M", "synthetic-8M"):
        return dict(nx=2000, ny=2000, ntree=400, reaches_per_tree=1000)  # Ne 8 000 000, Nr 400 000, Ns 1 200 000
    raise KeyError(size)


def _rng(seed, kind, row=0):
    return np.random.default_rng([seed, kind, row])


def make(nx, ny, ntree=None, reaches_per_tree=None, seed=SEED, shuffle=True, lake_frac=0.0, rows=None,
         stripe_rows=None):
    """rows=(r0, r1): build only the stripe of quad rows [r0, r1) (+ one halo row on each inner side) and
    return it as a partition (shud_up_b200.partition.extract) of the global nx x ny mesh; stripe_rows = rows
    per stripe of the global decomposition (owner rank of a row = row // stripe_rows)."""
    r0, r1 = (0, ny) if rows is None else rows
    qa, qb = max(r0 - 1, 0), min(r1 + 1, ny)       # quad rows generated (owned + halo candidates)
    nq = qb - qa
    # ---------------- nodes of rows qa .. qb ----------------
    jx = np.empty((nq + 1, nx + 1)); jy = np.empty((nq + 1, nx + 1))
    for j in range(nq + 1):
        u = _rng(seed, 1, qa + j).uniform(-0.2, 0.2, (2, nx + 1))
        jx[j], jy[j] = u[0], u[1]
    gx, gy = np.meshgrid(np.arange(nx + 1, dtype=np.float64), np.arange(qa, qb + 1, dtype=np.float64), indexing="xy")
    X = (gx + jx) * 100.0
    Yc = (gy + jy) * 100.0
    ph = _rng(seed, 0).uniform(0, 2 * np.pi, 6)
    noise = (np.sin(X / 3100.0 + ph[0]) * np.cos(Yc / 2300.0 + ph[1]) + 0.5 * np.sin(X / 900.0 + ph[2])
             * np.sin(Yc / 1300.0 + ph[3]) + 0.25 * np.cos(X / 410.0 + ph[4]) * np.cos(Yc / 370.0 + ph[5]))
    Z = 1000.0 + 0.01 * X + 0.005 * Yc + 5.0 * noise
    nid = lambda ix, iy: (iy - qa) * (nx + 1) + ix          # local node id of global grid point (ix, iy)
    Xf, Yf, Zf = X.ravel(), Yc.ravel(), Z.ravel()
    # ---------------- triangles: quad (ix,iy) -> cells 2q (touches the bottom edge) and 2q+1 (top edge) -----
    qx, qy = np.meshgrid(np.arange(nx), np.arange(qa, qb), indexing="xy")
    qx, qy = qx.ravel(), qy.ravel()
    n00, n10, n01, n11 = nid(qx, qy), nid(qx + 1, qy), nid(qx, qy + 1), nid(qx + 1, qy + 1)
    even = ((qx + qy) % 2) == 0
    # even: diagonal n00-n11 -> (n00,n10,n11) bottom, (n00,n11,n01) top ; odd: diagonal n10-n01
    tb = np.where(even[:, None], np.stack([n00, n10, n11], 1), np.stack([n00, n10, n01], 1))
    tt = np.where(even[:, None], np.stack([n00, n11, n01], 1), np.stack([n10, n11, n01], 1))
    Ne = 2 * nx * nq
    node = np.empty((Ne, 3), dtype=np.int64)
    node[0::2], node[1::2] = tb, tt
    gid = np.empty(Ne, dtype=np.int64)                      # global cell id
    gid[0::2] = 2 * (qy * nx + qx); gid[1::2] = 2 * (qy * nx + qx) + 1
    cell_row = np.repeat(qy, 2)
    x1, x2, x3 = Xf[node[:, 0]], Xf[node[:, 1]], Xf[node[:, 2]]
    y1, y2, y3 = Yf[node[:, 0]], Yf[node[:, 1]], Yf[node[:, 2]]
    area = 0.5 * ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1))
    assert (area > 0).all()
    cx, cy = (x1 + x2 + x3) / 3.0, (y1 + y2 + y3) / 3.0
    z_surf = (Zf[node[:, 0]] + Zf[node[:, 1]] + Zf[node[:, 2]]) / 3.0
    # edge j is opposite node j (edge[0] = node1-node2 in 0-based terms), Element.cpp:102-104
    ea = np.stack([node[:, 1], node[:, 2], node[:, 0]], 1)
    eb = np.stack([node[:, 2], node[:, 0], node[:, 1]], 1)
    ex1, ey1, ex2, ey2 = Xf[ea], Yf[ea], Xf[eb], Yf[eb]
    edge = np.hypot(ex2 - ex1, ey2 - ey1)
    # distance centroid -> edge line (foot of the perpendicular), Element.cpp:106-115
    dist2edge = np.abs((ex2 - ex1) * (ey1 - cy[:, None]) - (ex1 - cx[:, None]) * (ey2 - ey1)) / edge
    # neighbours through shared edges
    lo, hi = np.minimum(ea, eb).ravel(), np.maximum(ea, eb).ravel()
    key = lo * (Xf.size) + hi
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same = ks[1:] == ks[:-1]
    nabr = np.zeros(3 * Ne, dtype=np.int64)
    a, b = order[:-1][same], order[1:][same]
    nabr[a] = b // 3 + 1
    nabr[b] = a // 3 + 1
    nabr = nabr.reshape(Ne, 3)
    # ---------------- classes (global tables) and per-cell draws (per quad row) ----------------
    nsoil, ngeol, nlc = 16, 16, 12
    tr = _rng(seed, 3)
    U = lambda lo_, hi_, n: tr.uniform(lo_, hi_, n)
    soil = dict(infKsatV=U(1.0e-6, 4.8e-6, nsoil), ThetaS=U(0.39, 0.47, nsoil), ThetaR=np.full(nsoil, 0.01) + U(0, 0.03, nsoil),
                Alpha=U(2.6, 5.9, nsoil), Beta=U(1.13, 1.31, nsoil), hAreaF=U(0.005, 0.02, nsoil),
                macKsatV=U(0.010, 0.048, nsoil), infD=U(0.08, 0.15, nsoil))
    geol = dict(KsatH=U(7.2e-4, 2.4e-3, ngeol), KsatV=U(7.2e-5, 2.4e-4, ngeol), Sy=U(0.39, 0.47, ngeol),
                geo_vAreaF=U(0.005, 0.02, ngeol), macKsatH=U(7.2e-3, 2.4e-2, ngeol), macD=U(0.5, 2.0, ngeol))
    lc = dict(VegFrac=U(0.0, 0.75, nlc), Rough=U(5.8e-4, 7.5e-4, nlc), RzD=U(0.0, 0.6, nlc), SoilDgrd=U(0.0, 0.1, nlc),
              ImpAF=np.where(np.arange(nlc) % 4 == 0, U(0.0, 0.4, nlc), 0.0), lai=U(0.5, 3.4, nlc))
    NC = 2 * nx
    R = {k: np.empty(Ne) for k in ("aqd", "pe", "pt", "lai0", "prcp", "netf", "eic0", "eic1", "sf0", "sf1", "gw", "wet0",
                                   "wet1", "us")}
    isoil = np.empty(Ne, dtype=np.int64); igeol = np.empty(Ne, dtype=np.int64); ilc = np.empty(Ne, dtype=np.int64)
    for j in range(nq):
        g = _rng(seed, 2, qa + j)
        sl = slice(j * NC, (j + 1) * NC)
        isoil[sl], igeol[sl], ilc[sl] = g.integers(0, nsoil, NC), g.integers(0, ngeol, NC), g.integers(0, nlc, NC)
        for k in R:
            R[k][sl] = g.uniform(0, 1, NC)
    aqd = 10.0 + 20.0 * R["aqd"]
    m = {}
    m["ele_x"], m["ele_y"] = cx, cy
    m["ele_area"], m["ele_z_surf"], m["ele_z_bottom"] = area, z_surf, z_surf - aqd
    m["ele_depression"] = np.full(Ne, 0.0002)  # Element.hpp:93
    m["ele_AquiferDepth"] = m["ele_z_surf"] - m["ele_z_bottom"]  # InitElement, Element.cpp:229
    aqd = m["ele_AquiferDepth"]
    for k in ("ThetaS", "ThetaR", "Alpha", "Beta", "hAreaF", "infD"):
        m["ele_" + k] = soil[k][isoil]
    # initialize(): infKsatV, macKsatV *= 1-SoilDgrd; VegFrac *= 1-ImpAF (MD_initialize.cpp:184-186)
    m["ele_infKsatV"] = soil["infKsatV"][isoil] * (1 - lc["SoilDgrd"][ilc])
    m["ele_macKsatV"] = soil["macKsatV"][isoil] * (1 - lc["SoilDgrd"][ilc])
    m["ele_ThetaFC"] = m["ele_ThetaS"] * 0.75  # copySoil, Element.cpp:399
    for k in ("KsatH", "KsatV", "Sy", "geo_vAreaF", "macKsatH"):
        m["ele_" + k] = geol[k][igeol]
    m["ele_macD"] = np.minimum(geol["macD"][igeol], aqd)  # InitElement, Element.cpp:236-237
    m["ele_ImpAF"] = lc["ImpAF"][ilc]
    m["ele_VegFrac"] = lc["VegFrac"][ilc] * (1 - lc["ImpAF"][ilc])
    m["ele_Rough"] = lc["Rough"][ilc]
    m["ele_WetlandLevel"] = aqd - m["ele_infD"]
    m["ele_RootReachLevel"] = aqd - lc["RzD"][ilc]
    m["ele_QSS"] = np.zeros(Ne)
    has = nabr > 0
    nb0 = np.where(has, nabr - 1, 0)
    d2n = np.where(has, np.hypot(cx[:, None] - cx[nb0], cy[:, None] - cy[nb0]), 0.0)   # applyNabor, Element.cpp:256-266
    avgr = np.where(has, 0.5 * (m["ele_Rough"][:, None] + m["ele_Rough"][nb0]), m["ele_Rough"][:, None])
    m["ele_iBC"] = np.zeros(Ne, dtype=np.int32)
    m["ele_iSS"] = np.zeros(Ne, dtype=np.int32)
    # ---------------- lake (optional, whole-mesh builds only): a disc of cells ----------------
    ilake = np.zeros(Ne, dtype=np.int32)
    lakenabr = np.zeros((Ne, 3), dtype=np.int32)
    Nl = 0
    if lake_frac > 0:
        assert rows is None, "the lake variant is a single-partition mesh"
        r2 = lake_frac * (nx * 100.0) * (ny * 100.0) / np.pi
        inl = (cx - 0.62 * nx * 100.0) ** 2 + (cy - 0.5 * ny * 100.0) ** 2 < r2
        ilake[inl] = 1
        Nl = 1
        lakenabr = np.where(has & (ilake[nb0] > 0) & (ilake[:, None] <= 0), ilake[nb0], 0).astype(np.int32)  # MD_Lake.cpp:133-145
        zl = float(m["ele_z_surf"][inl].min()) - 2.0
        m["lake_zmin"] = np.array([zl])
        m["lake_NumEleLake"] = np.array([int(inl.sum())], dtype=np.int32)
        m["lake_bathy_ptr"] = np.array([0, 3], dtype=np.int32)
        A = float(area[inl].sum())
        m["lake_bathy_yi"] = np.array([zl, zl + 10.0, zl + 80.0])      # shape of input/qhh/qhh.lake.bathy
        m["lake_bathy_ai"] = np.array([0.92 * A, 0.92 * A, A])
    else:
        m["lake_zmin"] = np.zeros(0); m["lake_NumEleLake"] = np.zeros(0, dtype=np.int32)
        m["lake_bathy_ptr"] = np.zeros(1, dtype=np.int32); m["lake_bathy_yi"] = np.zeros(0); m["lake_bathy_ai"] = np.zeros(0)
    m["ele_iLake"] = ilake
    # ---------------- rivers: the trees whose band lies inside the owned rows ----------------
    if ntree is None:
        ntree = max(1, ny // 10)
    band = ny // ntree
    assert band >= 4, "need >= 4 quad rows per river tree"
    if reaches_per_tree is None:
        reaches_per_tree = min(nx, 1000)
    stem = int(0.6 * reaches_per_tree)
    trib = (reaches_per_tree - stem) // 4
    assert stem <= nx and trib >= 1
    per = stem + 4 * trib
    t_all = np.arange(ntree)
    t_idx = t_all[(t_all * band >= r0) & ((t_all + 1) * band <= r1)]
    nt = t_idx.size
    Nr = nt * per
    # river types: depth, bankslope, BottomWidth, rivRough, Cwr, KsatH, BedThick (input/ccw/ccw.sp.riv, 4 rows)
    rtype = np.array([[5.5, 0.0, 52.0, 0.04 / 60, 0.6, 0.1 / 1440, 0.1], [6.0, 0.5, 54.0, 0.04 / 60, 0.6, 0.1 / 1440, 0.1],
                      [6.5, 1.0, 56.0, 0.045 / 60, 0.6, 0.2 / 1440, 0.15], [7.0, 0.0, 58.0, 0.035 / 60, 0.62, 0.1 / 1440, 0.1]])
    r_ix = np.empty(Nr, dtype=np.int64); r_iy = np.empty(Nr, dtype=np.int64)
    r_down = np.empty(Nr, dtype=np.int64); r_type = np.empty(Nr, dtype=np.int64)
    iy0 = t_idx * band + band // 2
    # per tree: stem reaches ix = 0..stem-1 on line iy0 flowing to ix=0 (outlet, down=-3);
    # tributary j on line iy0+off_j covering ix = a_j .. a_j+trib-1, flowing to ix=a_j, joining the stem at ix=a_j
    offs = [1, -1, 1, -1]  # two tributaries per side line, on disjoint ix ranges
    starts = [int(stem * f) for f in (0.15, 0.35, 0.55, 0.75)]
    base = np.arange(nt) * per                      # first reach id (0-based, local) of each tree
    ixs = np.arange(stem)
    if nt:
        rid = (base[:, None] + ixs[None, :])        # stem reach ids
        r_ix[rid] = ixs[None, :]; r_iy[rid] = iy0[:, None]
        r_down[rid] = np.where(ixs[None, :] == 0, -3, rid)      # 1-based id of the reach at ix-1
        r_type[rid] = np.where(ixs[None, :] < stem // 3, 3, 2)
        for j in range(4):
            a = min(starts[j], nx - trib)
            tix = a + np.arange(trib)
            tid = base[:, None] + stem + j * trib + np.arange(trib)[None, :]
            r_ix[tid] = tix[None, :]; r_iy[tid] = (iy0 + offs[j])[:, None]
            dn = tid.copy()                         # 1-based id of the previous reach of the tributary
            dn[:, 0] = base + min(a, stem - 1) + 1  # joins the stem
            r_down[tid] = dn
            r_type[tid] = j % 2
    # geometry of a reach = the grid edge nodes (ix,iy)-(ix+1,iy)
    na, nb_ = nid(r_ix, r_iy), nid(r_ix + 1, r_iy)
    sinu = np.empty(Nr)
    for k, t in enumerate(t_idx):
        sinu[k * per:(k + 1) * per] = 1.0 + 0.2 * _rng(seed, 4, int(t)).uniform(0, 1, per)
    length = np.hypot(Xf[nb_] - Xf[na], Yf[nb_] - Yf[na]) * sinu
    zbed_a, zbed_b = Zf[na], Zf[nb_]
    slope = np.maximum(MINRIVSLOPE, (zbed_b - zbed_a) / length)   # flows towards decreasing x
    m["riv_Length"] = length
    m["riv_BedSlope"] = slope
    m["riv_depth"] = rtype[r_type, 0]; m["riv_bankslope"] = rtype[r_type, 1]; m["riv_BottomWidth"] = rtype[r_type, 2]
    m["riv_KsatH"] = rtype[r_type, 5]; m["riv_BedThick"] = rtype[r_type, 6]; m["riv_zbank"] = np.zeros(Nr)
    rr = rtype[r_type, 3]
    dn0 = np.where(r_down > 0, r_down - 1, 0)
    m["riv_avgRough"] = np.where(r_down > 0, 0.5 * (rr + rr[dn0]), rr)                 # River.cpp:74-90
    m["riv_Dist2DownStream"] = np.where(r_down > 0, 0.5 * (length + length[dn0]), length)
    m["riv_down"] = r_down.astype(np.int32)
    m["riv_BC"] = np.zeros(Nr, dtype=np.int32)
    m["riv_toLake"] = np.full(Nr, -9999, dtype=np.int32)
    # segments: 3 per reach - cell touching the edge from above (quad row iy, bottom triangle), from below
    # (quad row iy-1, top triangle) and the other triangle of the quad above
    q_up = (r_iy - qa) * nx + r_ix
    q_dn = (r_iy - 1 - qa) * nx + r_ix
    seg_ele = np.stack([2 * q_up, 2 * q_dn + 1, 2 * q_up + 1], 1).ravel() + 1
    seg_riv = np.repeat(np.arange(Nr) + 1, 3)
    seg_len = (np.stack([np.ones(Nr), np.ones(Nr), 0.3 * np.ones(Nr)], 1) * length[:, None]).ravel()
    if Nl:
        keep = ilake[seg_ele - 1] <= 0  # no river segments on lake cells
        seg_ele, seg_riv, seg_len = seg_ele[keep], seg_riv[keep], seg_len[keep]
    m["seg_iEle"] = seg_ele.astype(np.int32); m["seg_iRiv"] = seg_riv.astype(np.int32)
    m["seg_length"] = seg_len
    m["seg_Cwr"] = rtype[r_type[seg_riv - 1], 4]
    Ns = seg_ele.size
    # ---------------- one forcing step (a wet hour) ----------------
    m["qPotEvap"] = 2.4e-6 * (0.3 + 1.7 * R["pe"])
    m["qPotTran"] = 1.0e-6 * (0.5 + 3.0 * R["pt"])
    m["t_lai"] = lc["lai"][ilc] * np.where(R["lai0"] < 0.05, 0.0, 1.0)
    prcp = 2.0e-6 * (0.8 + 0.4 * R["prcp"])
    m["qElePrep"] = prcp
    m["qEleNetPrep"] = prcp * (0.7 + 0.3 * R["netf"])
    m["fu_Surf"] = np.ones(Ne); m["fu_Sub"] = np.ones(Ne)
    m["qEleE_IC_in"] = np.where(R["eic0"] < 0.3, 0.0, 1.5 * R["eic1"] * m["qPotTran"])
    # ---------------- state: every branch populated ----------------
    A = m["ele_AquiferDepth"]
    ysf = np.where(R["sf0"] < 0.5, 0.0, 0.02 * R["sf1"])
    ygw = (0.2 + 0.75 * R["gw"]) * A
    ygw = np.where(R["wet0"] < 0.04, (0.985 + 0.025 * R["wet1"]) * A, ygw)   # water table at / above the surface
    yus = (0.05 + 0.55 * R["us"]) * np.maximum(A - ygw, 0.02)
    yriv = np.empty(Nr)
    for k, t in enumerate(t_idx):
        yriv[k * per:(k + 1) * per] = _rng(seed, 5, int(t)).uniform(0, 0.5, per)
    yriv = yriv * m["riv_depth"]
    ylake = np.full(Nl, 8.0)
    m["y"] = np.concatenate([ysf, yus, ygw, yriv, ylake])
    # ---------------- random cell numbering (the locality pass has to undo it) ----------------
    if shuffle:
        p = _rng(seed, 6, r0).permutation(Ne)            # new id -> old id
        inv = np.empty(Ne, dtype=np.int64); inv[p] = np.arange(Ne)
        for kname in list(m.keys()):
            v = m[kname]
            if v.ndim == 1 and v.shape[0] == Ne and not kname.startswith(("riv_", "seg_", "lake_")) and kname != "y":
                m[kname] = v[p]
        m["y"] = np.concatenate([ysf[p], yus[p], ygw[p], yriv, ylake])
        nabr = np.where(nabr[p] > 0, inv[np.maximum(nabr[p] - 1, 0)] + 1, 0)
        edge, d2n, dist2edge, avgr, lakenabr = edge[p], d2n[p], dist2edge[p], avgr[p], lakenabr[p]
        gid, cell_row = gid[p], cell_row[p]
        m["seg_iEle"] = (inv[m["seg_iEle"] - 1] + 1).astype(np.int32)
    # [3][Ne] edge-major arrays
    m["ele_edge"] = np.ascontiguousarray(edge.T).ravel(); m["ele_Dist2Nabor"] = np.ascontiguousarray(d2n.T).ravel()
    m["ele_Dist2Edge"] = np.ascontiguousarray(dist2edge.T).ravel(); m["ele_avgRough"] = np.ascontiguousarray(avgr.T).ravel()
    m["ele_nabr"] = np.ascontiguousarray(nabr.T).ravel().astype(np.int32)
    m["ele_lakenabr"] = np.ascontiguousarray(lakenabr.T).ravel().astype(np.int32)
    for kname, v in (("Ne", Ne), ("Nr", Nr), ("Ns", Ns), ("Nl", Nl), ("close_boundary", 1), ("lakeon", 1 if Nl else 0)):
        m[kname] = np.array([v], dtype=np.int32)
    m["ele_gid"] = gid
    if rows is None:
        return m
    # ---------------- cut the owned rows out, keep the neighbours across the cut as halo ----------------
    from . import partition
    owned = (cell_row >= r0) & (cell_row < r1)
    part = cell_row // (stripe_rows or (r1 - r0))
    loc = partition.extract(m, owned, gid=gid, part_of_cell=part)
    return loc
