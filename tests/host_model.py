"""TEST INFRASTRUCTURE: a host N_Vector ops table (numpy, SUNDIALS-serial semantics) and a model adapter that
drives shud_up_b200.integrator / driver with the CPU oracle - the reference arm of the full-run comparison."""
import numpy as np

import oracle_lib


class HostOps:
    def N_VLinearSum(self, a, x, b, y, z): np.copyto(z, a * x + b * y)
    def N_VConst(self, c, z): z[:] = c
    def N_VProd(self, x, y, z): np.copyto(z, x * y)
    def N_VDiv(self, x, y, z): np.copyto(z, x / y)
    def N_VScale(self, c, x, z): np.copyto(z, c * x)
    def N_VAbs(self, x, z): np.copyto(z, np.abs(x))
    def N_VInv(self, x, z): np.copyto(z, 1.0 / x)
    def N_VAddConst(self, x, b, z): np.copyto(z, x + b)
    def N_VDotProd(self, x, y): return float(np.dot(x, y))
    def N_VWrmsNorm(self, x, w): return float(np.sqrt(np.sum((x * w) ** 2) / x.size))
    def N_VMaxNorm(self, x): return float(np.abs(x).max())
    def N_VMin(self, x): return float(x.min())
    def N_VL1Norm(self, x): return float(np.abs(x).sum())
    def N_VWSqrSumLocal(self, x, w): return float(np.sum((x * w) ** 2))
    def N_VDotProdMulti(self, x, Y): return np.array([float(np.dot(x, y)) for y in Y])
    def N_VWSqrSumMaskLocal(self, x, w, idv): return float(np.sum(((x * w) ** 2)[idv > 0]))
    def N_VWrmsNormMask(self, x, w, idv): return float(np.sqrt(self.N_VWSqrSumMaskLocal(x, w, idv) / x.size))
    def N_VMinQuotient(self, num, den): return float((num[den != 0] / den[den != 0]).min()) if (den != 0).any() else 1.7976931348623157e308

    def N_VInvTest(self, x, z):
        nz = x != 0
        z[nz] = 1.0 / x[nz]
        return bool(nz.all())

    def N_VConstrMask(self, c, x, m):
        bad = ((np.abs(c) > 1.5) & (x * c <= 0)) | ((np.abs(c) > 0.5) & (x * c < 0))
        m[:] = bad.astype(np.float64)
        return not bool(bad.any())

    def N_VLinearCombination(self, c, X, z):
        s = c[0] * X[0]
        for k in range(1, len(X)):
            s = s + c[k] * X[k]
        np.copyto(z, s)


class OracleModel:
    def __init__(self, mesh, fseq, land=None):
        """land: a --land-seq snapshot; the oracle's land-surface step then makes the forcing of each ET step"""
        self.mesh, self.fseq = dict(mesh), fseq
        self.land = None
        if land is not None:
            res = oracle_lib.oracle_land_seq(mesh, land)
            Ne = int(mesh["Ne"][0])
            self.land = {n: res[n] for n in ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "qElePrep", "fu_Surf", "fu_Sub",
                                             "qEleE_IC")}
        self.ops = HostOps()
        self.Ne, self.NY = int(mesh["Ne"][0]), int(np.asarray(mesh["y"]).size)
        self.satn = np.zeros(self.Ne)
        self.eic = np.zeros(self.Ne)
        self.outlets = np.nonzero(np.asarray(mesh["riv_down"]) < 0)[0]
        self.cur = dict(mesh)

    def new_vector(self):
        return np.zeros(self.NY)

    def set_forcing(self, k):
        if self.land is not None:
            for n in ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "qElePrep", "fu_Surf", "fu_Sub"):
                self.cur[n] = self.land[n][k]
            self.eic = np.array(self.land["qEleE_IC"][k], copy=True)
            return
        for n in ("qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "qElePrep"):
            self.cur[n] = self.fseq["fseq_" + n][k]
        self.cur["fu_Surf"] = np.ones(self.Ne); self.cur["fu_Sub"] = np.ones(self.Ne)
        self.eic = np.array(self.fseq["fseq_qEleE_IC"][k], copy=True)

    def _call(self, y, want_diag):
        out = oracle_lib.oracle_rhs(self.cur, y=y, u_satn=self.satn, qEleE_IC=self.eic, want_diag=want_diag)
        assert out["err"] == 0
        self.satn, self.eic = out["u_satn_out"], out["qEleE_IC_out"]
        return out

    def rhs(self, t, y, ydot):
        np.copyto(ydot, self._call(y, False)["ydot"])

    def load_state(self, y0):
        self.satn = np.zeros(self.Ne)
        return np.array(y0, dtype=np.float64, copy=True)

    def state_to_host(self, y):
        return np.array(y, copy=True)

    def outlet_flux(self, t, y):
        return self._call(y, True)["QrivDown"][self.outlets].copy()
