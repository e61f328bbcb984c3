"""TEST INFRASTRUCTURE: the checker arm of the integrator-level comparisons - oracle/_ref/libhostcv.so (host serial
N_Vector + oracle CVRhsFn + the SAME integrator source the product library compiles)."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib
from shud_up_b200 import abi, cvode

ROOT = oracle_lib.ROOT
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_ref", "libhostcv.so")
        srcs = [os.path.join(ROOT, "oracle", "host_cv.c"), os.path.join(ROOT, "shud_up_b200", "csrc", "shud_cvode.cpp"),
                os.path.join(ROOT, "oracle", "shud_oracle.c")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hostcv"], stdout=subprocess.DEVNULL)
        L = cvode.bind(C.CDLL(so))
        L.N_VNew_HostSerial.restype = C.c_void_p
        L.N_VNew_HostSerial.argtypes = [C.c_int64, C.c_void_p]
        L.N_VDisableFused_HostSerial.argtypes = [C.c_void_p]
        L.host_nv_opcount.restype = C.c_long
        _LIB = L
    return _LIB


class HostVector:
    def __init__(self, n, values=None):
        self.L = lib()
        self.h = C.c_void_p(self.L.N_VNew_HostSerial(int(n), None))
        self.n = int(n)
        if values is not None:
            self.array[:] = values

    @property
    def array(self):
        """numpy view of the vector's data (NV_DATA_S)"""
        return np.ctypeslib.as_array(self.L.N_VGetArrayPointer(self.h), shape=(self.n,))

    def close(self):
        if self.h:
            self.L.N_VDestroy(self.h)
            self.h = None


class OracleModelStruct(C.Structure):
    _fields_ = [("mesh", C.c_void_p), ("forcing", C.c_void_p), ("u_satn", C.POINTER(C.c_double)),
                ("qEleE_IC", C.POINTER(C.c_double)), ("nthreads", C.c_int), ("ncalls", C.c_long), ("last_err", C.c_int)]


class OracleCV:
    """the oracle RHS as a CVRhsFn with Model_Data's carried state, for shud_cv_create of libhostcv"""

    def __init__(self, snap):
        self.snap = dict(snap)
        self.mesh, self._keep_mesh = abi.make_mesh(self.snap)
        self.Ne = self.mesh.Ne
        self.satn = np.zeros(self.Ne)
        self.eic = np.zeros(self.Ne)
        self.M = OracleModelStruct()
        self.M.mesh = C.cast(C.pointer(self.mesh), C.c_void_p)
        self.M.u_satn = self.satn.ctypes.data_as(C.POINTER(C.c_double))
        self.M.qEleE_IC = self.eic.ctypes.data_as(C.POINTER(C.c_double))
        self.M.nthreads = 1
        self.set_forcing(self.snap, self.snap.get("qEleE_IC_in", np.zeros(self.Ne)))
        self.f_addr = cvode.fn_address(lib(), "shud_oracle_f")
        self.user_data = C.addressof(self.M)

    def set_forcing(self, arrays, qEleE_IC):
        """what updateforcing() + ET() leave for the following CVode() call (src/Model/shud.cpp:106-109)"""
        self.snap.update({k: np.ascontiguousarray(v, dtype=np.float64) for k, v in arrays.items()
                          if k in ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "qElePrep", "fu_Surf", "fu_Sub")})
        self.eic[:] = qEleE_IC
        self.forc, self._keep_forc = abi.make_forcing(self.snap, qEleE_IC=self.eic)
        self.M.forcing = C.cast(C.pointer(self.forc), C.c_void_p)

    def diag_call(self, y):
        """a direct f() with diagnostics at y (shud.cpp:127,141): moves the carried state on, like the reference's"""
        out = oracle_lib.oracle_rhs(self.snap, y=y, u_satn=self.satn, qEleE_IC=self.eic, want_diag=True)
        assert out["err"] == 0, out["err"]
        self.satn[:] = out["u_satn_out"]
        self.eic[:] = out["qEleE_IC_out"]
        return out


class OracleArm:
    """checker arm of shud_up_b200.driver.run_cv: oracle land-surface step + oracle RHS + host serial N_Vector under
    the integrator source of the product (libhostcv.so)"""

    def __init__(self, mesh, run):
        self.mesh = mesh
        self.lib = lib()
        self.model = OracleCV(mesh)
        self.Ne, self.NY = self.model.Ne, int(np.asarray(mesh["y"]).size)
        self.Nr = int(np.asarray(mesh["riv_down"]).size)
        self.yv = HostVector(self.NY, mesh["y"])
        self.y = self.yv.h
        self.f_addr, self.user_data = self.model.f_addr, self.model.user_data
        self.fused = None
        self.land, self._keep = abi.make_land(run)
        self.snow = np.array(run["land_yEleSnow0"], dtype=np.float64, copy=True)
        self.ics = np.array(run["land_yEleIS0"], dtype=np.float64, copy=True)
        self._steps = abi.land_steps(run)
        self._step_k = -1
        self.fn = oracle_lib.lib().shud_oracle_land_step
        self.fn.restype = C.c_int
        self.fn.argtypes = [C.c_void_p] * 3 + [C.POINTER(C.c_double)] * 2 + [C.c_void_p, C.POINTER(C.c_double)]
        self.cryo = None
        if self.land.cryosphere:
            oracle_lib.lib().shud_oracle_cryo_size.restype = C.c_long
            self.cryo = np.zeros(oracle_lib.lib().shud_oracle_cryo_size(C.byref(self.model.mesh), C.byref(self.land)))
            self.cryo[0] = -9999.0
        head_bc = np.asarray(mesh[abi._key(mesh, "iBC")]) > 0
        self._bc = (head_bc, np.asarray(mesh["riv_BC"]) > 0)

    def land_step(self, k):
        while self._step_k < k:
            self._step_k, S, keep = next(self._steps)
        o, arrs = abi.make_land_out(self.Ne)
        pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        rc = self.fn(C.byref(self.model.mesh), C.byref(self.land), C.byref(S), pd(self.snow), pd(self.ics), C.byref(o),
                     pd(self.cryo) if self.cryo is not None else None)
        assert rc == 0, rc
        self._lo = arrs
        self.model.set_forcing(arrs, arrs["qEleE_IC"])

    def land_out(self):
        out = dict(self._lo)
        out["yEleSnow"], out["yEleIS"] = self.snow.copy(), self.ics.copy()
        out["qEleE_IC"] = self.model.eic.copy()     # as the accepted-solution f() left it (same convention as the GPU arm)
        return out

    def state_host(self):
        """Model_Data::summary: BC heads / stages replace the solver's frozen rows (MD_update.cpp:190-216)"""
        y = self.yv.array.copy()
        Ne, Nr = self.Ne, self.Nr
        hb, rb = self._bc
        if hb.any():
            y[2 * Ne:3 * Ne][hb] = np.asarray(self.mesh["ele_yBC"])[hb]
        if rb.any():
            y[3 * Ne:3 * Ne + Nr][rb] = np.asarray(self.mesh["riv_yBC"])[rb]
        return y

    def diag(self, t):
        return self.model.diag_call(self.yv.array)

    def close(self):
        self.yv.close()
