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
    _fields_ = [("Nhalo", C.c_int32)] + [(n, _PD) for n in HALO_D]


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
