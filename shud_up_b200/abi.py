"""ctypes mirror of include/shud_b200.h (the C-ABI structs) and helpers that turn a
dict of numpy arrays (a snapshot, or a synthetic mesh) into those structs.

Field names are the reference's own (src/classes/Element.hpp, River.hpp, Lake.hpp);
snapshot keys carry an ``ele_`` / ``riv_`` / ``seg_`` / ``lake_`` prefix as written by
oracle/ref_driver.cpp.  No compute happens here.
"""
import ctypes as C

import numpy as np

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)

MESH_CELL_D = ["area", "z_surf", "z_bottom", "depression", "AquiferDepth", "Sy",
               "infD", "infKsatV", "macKsatV", "hAreaF", "ThetaS", "ThetaR", "ThetaFC", "Beta",
               "KsatH", "KsatV", "macKsatH", "macD", "geo_vAreaF",
               "VegFrac", "ImpAF", "WetlandLevel", "RootReachLevel", "Rough", "QSS"]
MESH_EDGE_D = ["edge", "Dist2Nabor", "Dist2Edge", "avgRough"]
MESH_EDGE_I = ["nabr", "lakenabr"]
MESH_CELL_I = ["iLake", "iBC", "iSS"]
MESH_RIV_D = ["riv_Length", "riv_BedSlope", "riv_depth", "riv_BottomWidth", "riv_bankslope",
              "riv_avgRough", "riv_Dist2DownStream", "riv_KsatH", "riv_BedThick", "riv_zbank"]
MESH_RIV_I = ["riv_down", "riv_BC", "riv_toLake"]
MESH_SEG_I = ["seg_iEle", "seg_iRiv"]
MESH_SEG_D = ["seg_length", "seg_Cwr"]


class ShudMesh(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("Ne", "Nr", "Ns", "Nl", "close_boundary", "lakeon")]
                + [(n, _PD) for n in MESH_CELL_D]
                + [(n, _PD) for n in MESH_EDGE_D]
                + [(n, _PI) for n in MESH_EDGE_I]
                + [(n, _PI) for n in MESH_CELL_I]
                + [("x", _PD), ("y", _PD)]
                + [(n, _PD) for n in MESH_RIV_D]
                + [(n, _PI) for n in MESH_RIV_I]
                + [(n, _PI) for n in MESH_SEG_I]
                + [(n, _PD) for n in MESH_SEG_D]
                + [("lake_zmin", _PD), ("lake_NumEleLake", _PI), ("lake_bathy_ptr", _PI),
                   ("lake_bathy_yi", _PD), ("lake_bathy_ai", _PD)])


HALO_D = ["z_surf", "z_bottom", "AquiferDepth", "macD", "macKsatH", "geo_vAreaF", "KsatH"]


class ShudHalo(C.Structure):
    _fields_ = [("Nhalo", C.c_int32)] + [(n, _PD) for n in HALO_D] + [("n_ghost_cells", C.c_int32), ("n_ghost_reaches", C.c_int32)]


def make_halo(halo):
    """dict with keys halo_<name> [Nhalo] -> (ShudHalo, keepalive)"""
    h = ShudHalo()
    keep = []
    n = int(np.asarray(halo["halo_z_surf"]).shape[0])
    h.Nhalo = n
    for name in HALO_D:
        a, p = _d(halo["halo_" + name])
        assert a.shape[0] == n
        keep.append(a)
        setattr(h, name, p)
    # cut river trees (partition.extract_cut): the last cells / reaches of the local mesh are ghosts
    h.n_ghost_cells = int(np.asarray(halo["n_ghost_cells"]).reshape(-1)[0]) if "n_ghost_cells" in halo else 0
    h.n_ghost_reaches = int(np.asarray(halo["n_ghost_reaches"]).reshape(-1)[0]) if "n_ghost_reaches" in halo else 0
    return h, keep


FORCING_D = ["qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "fu_Surf", "fu_Sub", "qElePrep",
             "qEleE_IC", "ele_yBC", "ele_QBC", "riv_yBC", "riv_qBC"]


class ShudForcing(C.Structure):
    _fields_ = [(n, _PD) for n in FORCING_D]


DIAG_CELL = ["qEleInfil", "qEleExfil", "qEleRecharge", "qEs", "qEu", "qEg", "qTu", "qTg",
             "qEleTrans", "qEleEvapo", "qEleETA", "iBeta", "u_effKH", "u_satn"]
DIAG_EDGE = ["QeleSurf", "QeleSub"]
DIAG_CELL2 = ["QeleSurfTot", "QeleSubTot", "Qe2r_Surf", "Qe2r_Sub"]
DIAG_SEG = ["QsegSurf", "QsegSub"]
DIAG_RIV = ["QrivSurf", "QrivSub", "QrivUp", "QrivDown"]
DIAG_LAKE = ["y2LakeArea", "QLakeSurf", "QLakeSub", "QLakeRivIn", "QLakeRivOut", "qLakeEvap", "qLakePrcp"]
DIAG_ALL = DIAG_CELL + DIAG_EDGE + DIAG_CELL2 + DIAG_SEG + DIAG_RIV + DIAG_LAKE


class ShudDiag(C.Structure):
    _fields_ = [(n, _PD) for n in DIAG_ALL]


LAND_OUT = ["qElePrep", "qPotEvap", "qPotTran", "qEleETP", "t_lai", "t_temp", "t_mf", "qEleNetPrep", "qEleE_IC", "fu_Surf",
            "fu_Sub", "rn_factor", "yEleSnow", "yEleIS"]


class ShudLand(C.Structure):  # include/shud_b200.h: shud_land
    _fields_ = ([("nforc", C.c_int32), ("nlc", C.c_int32), ("nmf", C.c_int32)]
                + [(n, _PI) for n in ("iForc", "iLC", "iMF")]
                + [(n, _PD) for n in ("Albedo", "FixPressure", "windH", "nx", "ny", "nz", "forc_z")]
                + [(n, C.c_double) for n in ("cPrep", "cTemp", "cLAItsd", "cMF", "cETP", "cISmax")]
                + [(n, C.c_int32) for n in ("radiation_is_net", "terrain_radiation", "cryosphere")]
                + [("rad_factor_cap", C.c_double), ("rad_cosz_min", C.c_double)]
                + [(n, C.c_double) for n in ("FT_surf_day", "FT_surf_max", "FT_surf_min", "FT_sub_day", "FT_sub_max",
                                             "FT_sub_min")])


class ShudLandStep(C.Structure):  # shud_land_step
    _fields_ = ([("forc", _PD), ("lai", _PD), ("mf", _PD), ("tsr_n", C.c_int32)]
                + [(n, _PD) for n in ("tsr_sx", "tsr_sy", "tsr_sz", "tsr_wdt")]
                + [("tsr_den", C.c_double), ("dt_min", C.c_double), ("t", C.c_double)])


class ShudLandOut(C.Structure):  # shud_land_out
    _fields_ = [(n, _PD) for n in LAND_OUT]


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_PD)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_PI)


def _key(snap, name):
    """snapshot key for a struct field (cell fields carry the ele_ prefix in snapshots)."""
    if name in snap:
        return name
    if "ele_" + name in snap:
        return "ele_" + name
    raise KeyError(name)


def make_mesh(snap):
    """dict of arrays -> (ShudMesh, keepalive list).  Accepts snapshot naming."""
    m = ShudMesh()
    keep = []
    for n in ("Ne", "Nr", "Ns", "Nl", "close_boundary", "lakeon"):
        setattr(m, n, int(np.asarray(snap[n]).reshape(-1)[0]))
    for names, conv in ((MESH_CELL_D + MESH_EDGE_D + MESH_RIV_D + MESH_SEG_D
                         + ["lake_zmin", "lake_bathy_yi", "lake_bathy_ai"], _d),
                        (MESH_EDGE_I + MESH_CELL_I + MESH_RIV_I + MESH_SEG_I
                         + ["lake_NumEleLake", "lake_bathy_ptr"], _i)):
        for n in names:
            a, p = conv(snap[_key(snap, n)])
            keep.append(a)
            setattr(m, n, p)
    for n in ("x", "y"):  # optional centroids ("y" alone is the state vector of a snapshot, never a centroid)
        src = snap.get("ele_" + n)
        if src is not None:
            a, p = _d(src)
            keep.append(a)
            setattr(m, n, p)
    return m, keep


def make_forcing(snap, qEleE_IC=None):
    f = ShudForcing()
    keep = []
    for n in FORCING_D:
        if n == "qEleE_IC":
            src = qEleE_IC if qEleE_IC is not None else snap.get("qEleE_IC_in", snap.get("qEleE_IC"))
        elif n in ("ele_yBC", "ele_QBC", "riv_yBC", "riv_qBC"):
            src = snap.get(n)
        else:
            src = snap[n]
        if src is None:
            setattr(f, n, None)
            continue
        a, p = _d(src)
        keep.append(a)
        setattr(f, n, p)
    return f, keep


def diag_sizes(Ne, Nr, Ns, Nl):
    sz = {}
    for n in DIAG_CELL + DIAG_CELL2:
        sz[n] = Ne
    for n in DIAG_EDGE:
        sz[n] = 3 * Ne
    for n in DIAG_SEG:
        sz[n] = Ns
    for n in DIAG_RIV:
        sz[n] = Nr
    for n in DIAG_LAKE:
        sz[n] = Nl
    return sz


def make_diag(Ne, Nr, Ns, Nl):
    """allocate every diag array; returns (ShudDiag, dict name->ndarray)."""
    d = ShudDiag()
    arrs = {}
    for n, k in diag_sizes(Ne, Nr, Ns, Nl).items():
        a = np.full(max(k, 1), np.nan, dtype=np.float64)[:k]
        arrs[n] = a
        setattr(d, n, a.ctypes.data_as(_PD) if k > 0 else None)
    return d, arrs


def make_land(snap):
    """land_* arrays of a --land-seq snapshot (oracle/ref_driver.cpp) -> (ShudLand, keepalive)"""
    L, keep = ShudLand(), []
    L.nforc, L.nlc, L.nmf = int(snap["land_nforc"][0]), int(snap["land_nlc"][0]), int(snap["land_nmf"][0])
    for n in ("iForc", "iLC", "iMF"):
        a, p = _i(snap["land_" + n]); keep.append(a); setattr(L, n, p)
    for n in ("Albedo", "FixPressure", "windH", "nx", "ny", "nz", "forc_z"):
        a, p = _d(snap["land_" + n]); keep.append(a); setattr(L, n, p)
    gc, cs = snap["land_gc"], snap["land_cs"]
    L.cPrep, L.cTemp, L.cLAItsd, L.cMF, L.cETP, L.cISmax = [float(v) for v in gc]
    L.radiation_is_net = int(cs[0] == cs[5])
    L.terrain_radiation, L.cryosphere = int(cs[1]), int(cs[2])
    L.rad_factor_cap, L.rad_cosz_min = float(cs[3]), float(cs[4])
    fz = snap["land_frozen"] if "land_frozen" in snap else [7., -1., -5., 28., -3., -10.]  # calib_frozen defaults
    L.FT_surf_day, L.FT_surf_max, L.FT_surf_min, L.FT_sub_day, L.FT_sub_max, L.FT_sub_min = [float(v) for v in fz]
    return L, keep


def land_steps(snap):
    """iterate the per-step inputs of a --land-seq snapshot: yields (k, ShudLandStep, keepalive)"""
    nf, nlc, nmf = int(snap["land_nforc"][0]), int(snap["land_nlc"][0]), int(snap["land_nmf"][0])
    n = np.asarray(snap["lseq_tsr_n"]); off = np.concatenate([[0], np.cumsum(n)])
    dt = float(snap["lseq_dt"][0]) if "lseq_dt" in snap else 60.0   # ET(t, tnext): tnext - t (MD_ET.cpp:286)
    for k in range(n.size):
        S, keep = ShudLandStep(), []
        for name, src in (("forc", snap["lseq_forc"][5 * nf * k:5 * nf * (k + 1)]), ("lai", snap["lseq_lai"][nlc * k:nlc * (k + 1)]),
                          ("mf", snap["lseq_mf"][nmf * k:nmf * (k + 1)])):
            a, p = _d(src); keep.append(a); setattr(S, name, p)
        for name in ("sx", "sy", "sz", "wdt"):
            a, p = _d(np.asarray(snap["lseq_tsr_" + name][off[k]:off[k + 1]])); keep.append(a); setattr(S, "tsr_" + name, p)
        S.tsr_n, S.tsr_den, S.dt_min, S.t = int(n[k]), float(snap["lseq_tsr_den"][k]), dt, float(snap["lseq_t"][k])
        yield k, S, keep


def make_land_out(Ne):
    o, arrs = ShudLandOut(), {}
    for n in LAND_OUT:
        a = np.full(Ne, np.nan); arrs[n] = a; setattr(o, n, a.ctypes.data_as(_PD))
    return o, arrs
