"""Loader for the CPU oracle (oracle/shud_oracle.c).  TEST INFRASTRUCTURE: importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_LIB = None
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from shud_up_b200 import abi, snapshot  # noqa: E402


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        L = C.CDLL(so)
        PD = C.POINTER(C.c_double)
        L.shud_oracle_rhs.restype = C.c_int
        L.shud_oracle_rhs.argtypes = [C.POINTER(abi.ShudMesh), C.POINTER(abi.ShudForcing), PD, PD, PD, PD,
                                      C.POINTER(abi.ShudDiag), C.c_int]
        L.shud_oracle_prime.restype = None
        L.shud_oracle_prime.argtypes = [C.POINTER(abi.ShudMesh), PD, PD]
        _LIB = L
    return _LIB


def load_case(basin, case):
    """static mesh of the basin overlaid with the case's arrays (mutations override)."""
    snap = snapshot.load(os.path.join(GOLDEN, f"{basin}.mesh.npz"))
    snap.update(snapshot.load(os.path.join(GOLDEN, f"{basin}.{case}.npz")))
    return snap


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def oracle_rhs(snap, y=None, u_satn=None, qEleE_IC=None, nthreads=1, want_diag=True):
    """one reference-equivalent f() call; returns dict(ydot, u_satn, qEleE_IC, err, diag arrays)."""
    mesh, keep = abi.make_mesh(snap)
    Ne, Nr, Ns, Nl = mesh.Ne, mesh.Nr, mesh.Ns, mesh.Nl
    y = np.ascontiguousarray(snap["y"] if y is None else y, dtype=np.float64)
    satn = np.array(snap["ele_u_satn"] if u_satn is None else u_satn, dtype=np.float64, copy=True)
    eic = np.array(snap["qEleE_IC_in"] if qEleE_IC is None else qEleE_IC, dtype=np.float64, copy=True)
    forc, keep2 = abi.make_forcing(snap, qEleE_IC=eic)
    ydot = np.empty_like(y)
    out = {}
    if want_diag:
        diag, arrs = abi.make_diag(Ne, Nr, Ns, Nl)
        dp = C.byref(diag)
    else:
        arrs, dp = {}, None
    err = lib().shud_oracle_rhs(C.byref(mesh), C.byref(forc), _pd(satn), _pd(eic), _pd(y), _pd(ydot), dp, nthreads)
    out.update(arrs)
    out.update(ydot=ydot, u_satn_out=satn, qEleE_IC_out=eic, err=err)
    return out


def oracle_prime(snap, y):
    mesh, keep = abi.make_mesh(snap)
    y = np.ascontiguousarray(y, dtype=np.float64)
    satn = np.empty(mesh.Ne, dtype=np.float64)
    lib().shud_oracle_prime(C.byref(mesh), _pd(y), _pd(satn))
    return satn


def oracle_land_seq(mesh_snap, land_snap):
    """replay every step of a --land-seq snapshot through the oracle; returns dict name -> [nstep, Ne]"""
    mesh, keep = abi.make_mesh(mesh_snap)
    L, keep2 = abi.make_land(land_snap)
    Ne = mesh.Ne
    snow = np.array(land_snap["land_yEleSnow0"], dtype=np.float64, copy=True)
    ics = np.array(land_snap["land_yEleIS0"], dtype=np.float64, copy=True)
    fn = lib().shud_oracle_land_step
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 3 + [C.POINTER(C.c_double)] * 2 + [C.c_void_p, C.POINTER(C.c_double)]
    cryo = None
    if L.cryosphere:
        lib().shud_oracle_cryo_size.restype = C.c_long
        cryo = np.zeros(lib().shud_oracle_cryo_size(C.byref(mesh), C.byref(L)))
        cryo[0] = -9999.0
    kept = set(int(v) for v in land_snap["lseq_kept"]) if "lseq_kept" in land_snap else None
    res = {n: [] for n in abi.LAND_OUT}
    for k, S, keep3 in abi.land_steps(land_snap):
        o, arrs = abi.make_land_out(Ne)
        rc = fn(C.byref(mesh), C.byref(L), C.byref(S), _pd(snow), _pd(ics), C.byref(o), _pd(cryo) if cryo is not None else None)
        assert rc == 0, rc
        if kept is not None and k not in kept:
            continue
        for n in abi.LAND_OUT:
            res[n].append(arrs[n].copy())
    return {n: np.stack(v) for n, v in res.items()}
