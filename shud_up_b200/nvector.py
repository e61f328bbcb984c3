"""Device N_Vector: host-side mirror of the SUNDIALS N_Vector operations CVODE/SPGMR call on the
SHUD state vectors (reference: N_VNew_Serial / N_VNew_OpenMP at src/Model/shud.cpp:59-64; ops table
per SUNDIALS 6, see include/shud_nvector.h).  Function names follow SUNDIALS (N_VLinearSum, ...);
vectors are torch CUDA float64 tensors (device memory plumbing only) and every operation is a call
into libshud_b200.so - nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import api

_PD = C.POINTER(C.c_double)
_sigs_done = False
MAXVEC = 8


def _lib():
    global _sigs_done
    L = api.lib()
    if not _sigs_done:
        vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
        PP = C.POINTER(C.c_void_p)
        sig = {
            "shud_nv_ws_create": [ci, vp, C.POINTER(vp)],
            "shud_nv_ws_set_peer_allreduce": [vp, ci, ci, C.POINTER(vp)],
            "shud_nv_ws_local": [vp, ci],
            "shud_nv_linearsum": [vp, i64, dbl, vp, dbl, vp, vp],
            "shud_nv_const": [vp, i64, dbl, vp],
            "shud_nv_prod": [vp, i64, vp, vp, vp],
            "shud_nv_div": [vp, i64, vp, vp, vp],
            "shud_nv_scale": [vp, i64, dbl, vp, vp],
            "shud_nv_abs": [vp, i64, vp, vp],
            "shud_nv_inv": [vp, i64, vp, vp],
            "shud_nv_addconst": [vp, i64, vp, dbl, vp],
            "shud_nv_compare": [vp, i64, dbl, vp, vp],
            "shud_nv_dotprod": [vp, i64, vp, vp, _PD],
            "shud_nv_maxnorm": [vp, i64, vp, _PD],
            "shud_nv_min": [vp, i64, vp, _PD],
            "shud_nv_l1norm": [vp, i64, vp, _PD],
            "shud_nv_wsqrsum": [vp, i64, vp, vp, _PD],
            "shud_nv_wsqrsum_mask": [vp, i64, vp, vp, vp, _PD],
            "shud_nv_wrmsnorm": [vp, i64, vp, vp, i64, _PD],
            "shud_nv_wrmsnorm_mask": [vp, i64, vp, vp, vp, i64, _PD],
            "shud_nv_wl2norm": [vp, i64, vp, vp, _PD],
            "shud_nv_invtest": [vp, i64, vp, vp, C.POINTER(ci)],
            "shud_nv_constrmask": [vp, i64, vp, vp, vp, C.POINTER(ci)],
            "shud_nv_minquotient": [vp, i64, vp, vp, _PD],
            "shud_nv_linearcombination": [vp, i64, ci, _PD, PP, vp],
            "shud_nv_scaleaddmulti": [vp, i64, ci, _PD, vp, PP, PP],
            "shud_nv_dotprodmulti": [vp, i64, ci, vp, PP, _PD],
            "shud_nv_linearsumvectorarray": [vp, i64, ci, dbl, PP, dbl, PP, PP],
            "shud_nv_scalevectorarray": [vp, i64, ci, _PD, PP, PP],
            "shud_nv_constvectorarray": [vp, i64, ci, dbl, PP],
            "shud_nv_wrmsnormvectorarray": [vp, i64, ci, PP, PP, i64, _PD],
            "shud_spgmr_create": [vp, vp, ci, i64, C.POINTER(vp)],
            "shud_spgmr_solve": [vp, dbl, dbl, vp, vp, vp, vp, dbl, vp, C.POINTER(ci), _PD],
            "shud_nv_ewt": [vp, i64, dbl, dbl, vp, vp],
            "shud_nv_newton_resid": [vp, i64, dbl, vp, vp, vp, vp],
            "shud_nv_newton_update": [vp, i64, vp, vp, i64, vp, vp, _PD],
            "shud_nv_dq_perturb": [vp, i64, dbl, vp, vp, vp, vp],
            "shud_nv_dq_combine": [vp, i64, dbl, dbl, vp, vp, vp, vp, vp],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = args
        L.shud_nv_ws_destroy.restype = None
        L.shud_nv_ws_local.restype = None
        L.shud_nv_ws_destroy.argtypes = [vp]
        L.shud_spgmr_destroy.restype = None
        L.shud_spgmr_destroy.argtypes = [vp]
        _sigs_done = True
    return L


def _p(t):
    return C.c_void_p(t.data_ptr())


def _pp(ts):
    arr = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    return arr


def _coef(c):
    a = np.ascontiguousarray(c, dtype=np.float64)
    return a, a.ctypes.data_as(_PD)


class NVectorOps:
    """The ops table, bound to one device and one CUDA stream (a cudaStream_t handle or 0)."""

    def __init__(self, device=0, stream_ptr=None, n_global=None, owner=None):
        """owner: the object that owns the stream (kept alive for as long as this table exists)"""
        self._owner = owner
        L = _lib()
        h = C.c_void_p()
        api._chk(L.shud_nv_ws_create(int(device), C.c_void_p(stream_ptr or 0), C.byref(h)), "shud_nv_ws_create")
        self._h, self._L = h, L
        self.n_global = n_global  # length of the whole distributed vector (None: local length)

    def close(self):
        if getattr(self, "_h", None):
            self._L.shud_nv_ws_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _c(self, rc, what):
        api._chk(rc, what)

    def set_peer_allreduce(self, nranks, rank, boxes):
        """every reduction of this table becomes a global one: the reduction kernels combine the ranks' partial results
        through the mailboxes `boxes` (ShudRHS.p2p_mailboxes) over NVLink; nranks <= 1 switches it off"""
        arr = (C.c_void_p * 16)(*([b for b in boxes] + [None] * (16 - len(boxes))))
        self._c(self._L.shud_nv_ws_set_peer_allreduce(self._h, int(nranks), int(rank), arr), "set_peer_allreduce")

    def local(self, on):
        """reductions stay local while on (the *local members of the operations table)"""
        self._L.shud_nv_ws_local(self._h, 1 if on else 0)

    # ---- streaming ----
    def N_VLinearSum(self, a, x, b, y, z):
        self._c(self._L.shud_nv_linearsum(self._h, x.numel(), a, _p(x), b, _p(y), _p(z)), "N_VLinearSum")

    def N_VConst(self, c, z):
        self._c(self._L.shud_nv_const(self._h, z.numel(), c, _p(z)), "N_VConst")

    def N_VProd(self, x, y, z):
        self._c(self._L.shud_nv_prod(self._h, x.numel(), _p(x), _p(y), _p(z)), "N_VProd")

    def N_VDiv(self, x, y, z):
        self._c(self._L.shud_nv_div(self._h, x.numel(), _p(x), _p(y), _p(z)), "N_VDiv")

    def N_VScale(self, c, x, z):
        self._c(self._L.shud_nv_scale(self._h, x.numel(), c, _p(x), _p(z)), "N_VScale")

    def N_VAbs(self, x, z):
        self._c(self._L.shud_nv_abs(self._h, x.numel(), _p(x), _p(z)), "N_VAbs")

    def N_VInv(self, x, z):
        self._c(self._L.shud_nv_inv(self._h, x.numel(), _p(x), _p(z)), "N_VInv")

    def N_VAddConst(self, x, b, z):
        self._c(self._L.shud_nv_addconst(self._h, x.numel(), _p(x), b, _p(z)), "N_VAddConst")

    def N_VCompare(self, c, x, z):
        self._c(self._L.shud_nv_compare(self._h, x.numel(), c, _p(x), _p(z)), "N_VCompare")

    # ---- reductions (local part; a distributed caller combines with an allreduce) ----
    def _red(self, fn, *args):
        out = C.c_double(0.0)
        self._c(fn(self._h, *args, C.byref(out)), fn.__name__)
        return out.value

    def N_VDotProd(self, x, y):
        return self._red(self._L.shud_nv_dotprod, x.numel(), _p(x), _p(y))

    def N_VMaxNorm(self, x):
        return self._red(self._L.shud_nv_maxnorm, x.numel(), _p(x))

    def N_VMin(self, x):
        return self._red(self._L.shud_nv_min, x.numel(), _p(x))

    def N_VL1Norm(self, x):
        return self._red(self._L.shud_nv_l1norm, x.numel(), _p(x))

    def N_VWSqrSumLocal(self, x, w):
        return self._red(self._L.shud_nv_wsqrsum, x.numel(), _p(x), _p(w))

    def N_VWSqrSumMaskLocal(self, x, w, idv):
        return self._red(self._L.shud_nv_wsqrsum_mask, x.numel(), _p(x), _p(w), _p(idv))

    def N_VWrmsNorm(self, x, w):
        return self._red(self._L.shud_nv_wrmsnorm, x.numel(), _p(x), _p(w), int(self.n_global or x.numel()))

    def N_VWrmsNormMask(self, x, w, idv):
        return self._red(self._L.shud_nv_wrmsnorm_mask, x.numel(), _p(x), _p(w), _p(idv), int(self.n_global or x.numel()))

    def N_VWL2Norm(self, x, w):
        return self._red(self._L.shud_nv_wl2norm, x.numel(), _p(x), _p(w))

    def N_VMinQuotient(self, num, den):
        return self._red(self._L.shud_nv_minquotient, num.numel(), _p(num), _p(den))

    def N_VInvTest(self, x, z):
        ok = C.c_int(0)
        self._c(self._L.shud_nv_invtest(self._h, x.numel(), _p(x), _p(z), C.byref(ok)), "N_VInvTest")
        return bool(ok.value)

    def N_VConstrMask(self, c, x, m):
        ok = C.c_int(0)
        self._c(self._L.shud_nv_constrmask(self._h, x.numel(), _p(c), _p(x), _p(m), C.byref(ok)), "N_VConstrMask")
        return bool(ok.value)

    # ---- fused ----
    def N_VLinearCombination(self, c, X, z):
        a, pa = _coef(c)
        self._c(self._L.shud_nv_linearcombination(self._h, z.numel(), len(X), pa, _pp(X), _p(z)), "N_VLinearCombination")

    def N_VScaleAddMulti(self, a, x, Y, Z):
        aa, pa = _coef(a)
        self._c(self._L.shud_nv_scaleaddmulti(self._h, x.numel(), len(Y), pa, _p(x), _pp(Y), _pp(Z)), "N_VScaleAddMulti")

    def N_VDotProdMulti(self, x, Y):
        out = np.zeros(len(Y))
        self._c(self._L.shud_nv_dotprodmulti(self._h, x.numel(), len(Y), _p(x), _pp(Y), out.ctypes.data_as(_PD)),
                "N_VDotProdMulti")
        return out

    def N_VLinearSumVectorArray(self, a, X, b, Y, Z):
        self._c(self._L.shud_nv_linearsumvectorarray(self._h, X[0].numel(), len(X), a, _pp(X), b, _pp(Y), _pp(Z)),
                "N_VLinearSumVectorArray")

    def N_VScaleVectorArray(self, c, X, Z):
        a, pa = _coef(c)
        self._c(self._L.shud_nv_scalevectorarray(self._h, X[0].numel(), len(X), pa, _pp(X), _pp(Z)), "N_VScaleVectorArray")

    def N_VConstVectorArray(self, c, Z):
        self._c(self._L.shud_nv_constvectorarray(self._h, Z[0].numel(), len(Z), c, _pp(Z)), "N_VConstVectorArray")

    def N_VWrmsNormVectorArray(self, X, W):
        out = np.zeros(len(X))
        self._c(self._L.shud_nv_wrmsnormvectorarray(self._h, X[0].numel(), len(X), _pp(X), _pp(W),
                                                    int(self.n_global or X[0].numel()), out.ctypes.data_as(_PD)),
                "N_VWrmsNormVectorArray")
        return out

    # ---- integrator-level fusions ----
    def EwtSet(self, rtol, atol, y, ewt):
        self._c(self._L.shud_nv_ewt(self._h, y.numel(), rtol, atol, _p(y), _p(ewt)), "EwtSet")

    def NewtonResid(self, gamma, f, psi, y, r):
        self._c(self._L.shud_nv_newton_resid(self._h, y.numel(), gamma, _p(f), _p(psi), _p(y), _p(r)), "NewtonResid")

    def NewtonUpdate(self, x, ewt, y, acor):
        out = C.c_double(0.0)
        self._c(self._L.shud_nv_newton_update(self._h, x.numel(), _p(x), _p(ewt), int(self.n_global or x.numel()), _p(y),
                                              _p(acor), C.byref(out)), "NewtonUpdate")
        return out.value

    def DQPerturb(self, sigma, vs, ewt, y, ytemp):
        self._c(self._L.shud_nv_dq_perturb(self._h, y.numel(), sigma, _p(vs), _p(ewt), _p(y), _p(ytemp)), "DQPerturb")

    def DQCombine(self, sigma, gamma, vs, ewt, fpert, fy, out):
        self._c(self._L.shud_nv_dq_combine(self._h, fy.numel(), sigma, gamma, _p(vs), _p(ewt), _p(fpert), _p(fy), _p(out)),
                "DQCombine")


class DeviceSPGMR:
    """SUNLinSol_SPGMR + CVLS difference-quotient Jv, resident on the device (shud_spgmr_* of the C ABI)."""

    def __init__(self, ops, shud_rhs, maxl=5, n_global=0):
        self._ops, self._rhs, self._L = ops, shud_rhs, _lib()
        h = C.c_void_p()
        api._chk(self._L.shud_spgmr_create(shud_rhs._h, ops._h, int(maxl), int(n_global), C.byref(h)), "shud_spgmr_create")
        self._h = h

    def solve(self, t, gamma, y, fy, ewt, b, tol, x):
        nli, res = C.c_int(0), C.c_double(0.0)
        rc = self._L.shud_spgmr_solve(self._h, float(t), float(gamma), _p(y), _p(fy), _p(ewt), _p(b), float(tol), _p(x),
                                      C.byref(nli), C.byref(res))
        if rc < 0:
            api._chk(rc, "shud_spgmr_solve")
        return rc, nli.value, res.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.shud_spgmr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistributedOps:
    """N_Vector ops of a vector distributed over the ranks of a torch.distributed group (one partition per GPU):
    streaming ops are local; every reduction is the local kernel followed by an allreduce of the scalar(s)
    (SUNDIALS' nvdotprodlocal / nvwsqrsumlocal / nvmaxnormlocal / nvminlocal + nvgetcommunicator pattern).
    `local` is an NVectorOps (or any table with the same N_V* methods); n_global = sum of the local lengths."""

    def __init__(self, local, dist, device, n_global):
        import torch
        self._l, self._dist, self._torch, self._dev, self.n_global = local, dist, torch, device, int(n_global)

    _LOCAL_OK = ("N_VLinearSum", "N_VConst", "N_VProd", "N_VDiv", "N_VScale", "N_VAbs", "N_VInv", "N_VAddConst",
                 "N_VCompare", "N_VLinearCombination", "N_VScaleAddMulti", "N_VLinearSumVectorArray",
                 "N_VScaleVectorArray", "N_VConstVectorArray", "EwtSet", "NewtonResid", "DQPerturb", "DQCombine",
                 "close", "sync")

    def __getattr__(self, name):          # streaming and fused streaming ops only: purely local, no reduction inside
        if name in DistributedOps._LOCAL_OK:
            return getattr(self._l, name)
        raise AttributeError(f"DistributedOps has no op {name} (reductions must be overridden with an allreduce)")

    def _allreduce(self, vals, op):
        t = self._torch.tensor(vals, dtype=self._torch.float64, device=self._dev)
        self._dist.all_reduce(t, op=op)
        return t.tolist()

    def N_VDotProd(self, x, y):
        return self._allreduce([self._l.N_VDotProd(x, y)], self._dist.ReduceOp.SUM)[0]

    def N_VDotProdMulti(self, x, Y):
        import numpy as np
        return np.array(self._allreduce(list(self._l.N_VDotProdMulti(x, Y)), self._dist.ReduceOp.SUM))

    def N_VMaxNorm(self, x):
        return self._allreduce([self._l.N_VMaxNorm(x)], self._dist.ReduceOp.MAX)[0]

    def N_VMin(self, x):
        return self._allreduce([self._l.N_VMin(x)], self._dist.ReduceOp.MIN)[0]

    def N_VL1Norm(self, x):
        return self._allreduce([self._l.N_VL1Norm(x)], self._dist.ReduceOp.SUM)[0]

    def N_VWSqrSumLocal(self, x, w):
        return self._l.N_VWSqrSumLocal(x, w)

    def N_VWrmsNorm(self, x, w):
        s = self._allreduce([self._l.N_VWSqrSumLocal(x, w)], self._dist.ReduceOp.SUM)[0]
        return (s / self.n_global) ** 0.5

    def N_VWL2Norm(self, x, w):
        return self._allreduce([self._l.N_VWSqrSumLocal(x, w)], self._dist.ReduceOp.SUM)[0] ** 0.5

    def N_VWSqrSumMaskLocal(self, x, w, idv):
        return self._l.N_VWSqrSumMaskLocal(x, w, idv)

    def N_VWrmsNormMask(self, x, w, idv):
        s = self._allreduce([self._l.N_VWSqrSumMaskLocal(x, w, idv)], self._dist.ReduceOp.SUM)[0]
        return (s / self.n_global) ** 0.5

    def N_VMinQuotient(self, num, den):
        return self._allreduce([self._l.N_VMinQuotient(num, den)], self._dist.ReduceOp.MIN)[0]

    def N_VInvTest(self, x, z):
        return bool(self._allreduce([1.0 if self._l.N_VInvTest(x, z) else 0.0], self._dist.ReduceOp.MIN)[0])

    def N_VConstrMask(self, c, x, m):
        return bool(self._allreduce([1.0 if self._l.N_VConstrMask(c, x, m) else 0.0], self._dist.ReduceOp.MIN)[0])

    def N_VWrmsNormVectorArray(self, X, W):
        import numpy as np
        s = self._allreduce([self._l.N_VWSqrSumLocal(x, w) for x, w in zip(X, W)], self._dist.ReduceOp.SUM)
        return np.sqrt(np.array(s) / self.n_global)

    def NewtonUpdate(self, x, ewt, y, acor):
        """y += x; acor += x locally, ||x||_WRMS over the WHOLE vector: every rank gets the same convergence norm
        (the local table would return sqrt(local sum / n_global))"""
        s = self._l.N_VWSqrSumLocal(x, ewt)
        self._l.N_VLinearSum(1.0, y, 1.0, x, y)
        self._l.N_VLinearSum(1.0, acor, 1.0, x, acor)
        return (self._allreduce([s], self._dist.ReduceOp.SUM)[0] / self.n_global) ** 0.5
