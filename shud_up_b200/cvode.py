"""ctypes mirror of include/shud_cvode.h + include/shud_sundials.h: the CVODE-shaped integrator (C,
shud_up_b200/csrc/shud_cvode.cpp) on SUNDIALS-6-layout N_Vectors.  The Python side only wires pointers together: the
time loop, the Newton-Krylov iteration and every vector operation run in the library.

Mirrors the calls of the reference's driver (src/Model/shud.cpp:59-64,78,131; src/Equations/cvode_config.cpp:162-193):
    udata = N_VNew_ShudB200(NY, ws, gpu)                 N_VNew_Serial(NY, sunctx)
    cv = CVode(lib, f, gpu, t0, udata); cv.configure()   SetCVODE(mem, f, MD, udata, LS, sunctx)
    cv.solve(tout, udata)                                CVode(mem, tout, udata, &t, CV_NORMAL)
"""
import ctypes as C

CV_NORMAL, CV_ONE_STEP = 1, 2
CV_SUCCESS, CV_TSTOP_RETURN = 0, 1


class Stats(C.Structure):
    _fields_ = [(n, C.c_long) for n in ("nst", "nfe", "nfeLS", "nni", "nli", "ncfn", "netf", "ncfl")] + \
               [("qlast", C.c_int), ("qcur", C.c_int)] + \
               [(n, C.c_double) for n in ("hinused", "hlast", "hcur", "tcur")]


class Fused(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("ewt_set", C.c_void_p), ("nls_residual", C.c_void_p), ("lsolve", C.c_void_p),
                ("predict", C.c_void_p), ("newton_step", C.c_void_p), ("ewt_set_norm", C.c_void_p),
                ("complete_step", C.c_void_p)]


def bind(lib):
    """declare the shud_cv_* entry points on a loaded library (the product library or the CPU checker library)"""
    if getattr(lib, "_shud_cv_bound", False):
        return lib
    vp, d = C.c_void_p, C.c_double
    sig = {
        "shud_cv_create": (C.c_int, [vp, vp, d, vp, C.POINTER(vp)]),
        "shud_cv_free": (None, [vp]),
        "shud_cv_reinit": (C.c_int, [vp, d, vp]),
        "shud_cv_sstolerances": (C.c_int, [vp, d, d]),
        "shud_cv_set_max_ord": (C.c_int, [vp, C.c_int]),
        "shud_cv_set_min_step": (C.c_int, [vp, d]),
        "shud_cv_set_max_step": (C.c_int, [vp, d]),
        "shud_cv_set_init_step": (C.c_int, [vp, d]),
        "shud_cv_set_max_num_steps": (C.c_int, [vp, C.c_long]),
        "shud_cv_set_stop_time": (C.c_int, [vp, d]),
        "shud_cv_set_maxl": (C.c_int, [vp, C.c_int]),
        "shud_cv_set_fused": (C.c_int, [vp, C.POINTER(Fused)]),
        "shud_cv_solve": (C.c_int, [vp, d, vp, C.POINTER(d), C.c_int]),
        "shud_cv_get_dky": (C.c_int, [vp, d, C.c_int, vp]),
        "shud_cv_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
        "shud_cv_linsolve": (C.c_int, [vp, d, d, vp, vp, vp, vp, d, vp, C.POINTER(C.c_int)]),
        "N_VAbs": (None, [vp, vp]), "N_VInv": (None, [vp, vp]), "N_VAddConst": (None, [vp, d, vp]),
        "N_VClone": (vp, [vp]),
        "N_VDestroy": (None, [vp]),
        "N_VGetArrayPointer": (C.POINTER(d), [vp]),
        "N_VGetLength": (C.c_int64, [vp]),
        "N_VLinearSum": (None, [d, vp, d, vp, vp]),
        "N_VWrmsNorm": (d, [vp, vp]),
        "N_VDotProd": (d, [vp, vp]),
        "N_VMaxNorm": (d, [vp]),
        "N_VMin": (d, [vp]),
        "N_VConst": (None, [d, vp]),
        "N_VScale": (None, [d, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib._shud_cv_bound = True
    return lib


def fn_address(lib, name):
    """address of an exported function (to hand a CVRhsFn of the library to shud_cv_create)"""
    return C.cast(getattr(lib, name), C.c_void_p).value


class CVError(RuntimeError):
    def __init__(self, flag, what):
        super().__init__(f"{what}: CVODE flag {flag}")
        self.flag = flag


class CVode:
    """CVodeCreate(CV_BDF) + CVodeInit + CVodeSetUserData on an N_Vector handle"""

    def __init__(self, lib, f_addr, user_data, t0, y0):
        self.lib = bind(lib)
        h = C.c_void_p()
        rc = lib.shud_cv_create(C.c_void_p(f_addr), C.c_void_p(user_data), float(t0), y0, C.byref(h))
        if rc:
            raise CVError(rc, "shud_cv_create")
        self._h = h
        self.t = float(t0)

    def configure(self, rtol=1e-4, atol=1e-4, init_step=0.0, max_step=0.0, min_step=1e-6, max_num_steps=1000000, maxl=0,
                  max_ord=5):
        """SetCVODE (cvode_config.cpp:162-193): SStolerances, SPGMR(PREC_NONE, maxl), Init/Min/MaxStep, MaxNumSteps"""
        L, h = self.lib, self._h
        for rc, what in ((L.shud_cv_sstolerances(h, rtol, atol), "sstolerances"), (L.shud_cv_set_maxl(h, maxl), "maxl"),
                         (L.shud_cv_set_init_step(h, init_step), "init_step"), (L.shud_cv_set_min_step(h, min_step), "min_step"),
                         (L.shud_cv_set_max_step(h, max_step), "max_step"),
                         (L.shud_cv_set_max_num_steps(h, max_num_steps), "max_num_steps"),
                         (L.shud_cv_set_max_ord(h, max_ord), "max_ord")):
            if rc:
                raise CVError(rc, what)
        return self

    def set_fused(self, fused):
        self._fused = fused  # keep alive
        rc = self.lib.shud_cv_set_fused(self._h, C.byref(fused) if fused is not None else None)
        if rc:
            raise CVError(rc, "set_fused")

    def set_stop_time(self, tstop):
        rc = self.lib.shud_cv_set_stop_time(self._h, float(tstop))
        if rc:
            raise CVError(rc, "set_stop_time")

    def solve(self, tout, yout, itask=CV_NORMAL):
        t = C.c_double(self.t)
        rc = self.lib.shud_cv_solve(self._h, float(tout), yout, C.byref(t), itask)
        self.t = t.value
        if rc < 0:
            raise CVError(rc, f"shud_cv_solve(tout={tout}) stopped at t={t.value}")
        return rc

    def get_dky(self, t, k, dky):
        rc = self.lib.shud_cv_get_dky(self._h, float(t), int(k), dky)
        if rc:
            raise CVError(rc, "get_dky")

    def linsolve(self, t, gamma, y, fy, ewt, b, delta, x):
        """one SPGMR solve of (I - gamma J) x = b as CVLS drives it; returns (code, Krylov iterations)"""
        nli = C.c_int(0)
        rc = self.lib.shud_cv_linsolve(self._h, float(t), float(gamma), y, fy, ewt, b, float(delta), x, C.byref(nli))
        if rc < 0:
            raise CVError(rc, "shud_cv_linsolve")
        return rc, nli.value

    def stats(self):
        s = Stats()
        self.lib.shud_cv_get_stats(self._h, C.byref(s))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def close(self):
        if self._h:
            self.lib.shud_cv_free(self._h)
            self._h = None
