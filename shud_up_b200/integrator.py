"""A CVODE-shaped implicit integrator written against the N_Vector ops table, so that the same code drives the
GPU path (NVectorOps + ShudRHS.f_dev) and, in tests, a host ops table + the CPU oracle.

SUNDIALS is not vendored in the reference and not installed here, so full runs cannot be
made with CVODE itself.  This module follows the structure and the settings the reference configures
(src/Equations/cvode_config.cpp:162-193, src/Model/shud.cpp:89-133):
  * BDF, variable step, variable order 1..max_order (default 2), scalar rtol/atol, error weights
    ewt = 1/(rtol |y| + atol), WRMS norms, MaxStep / InitStep / MinStep from cfg.para;
  * modified Newton-Krylov: <= 3 Newton iterations, convergence test del*min(1,crate) <= 0.1 (nlscoef);
  * SPGMR: maxl = 5, no preconditioner, modified Gram-Schmidt, no restarts, eps_lin = 0.05, ewt scaling on
    both sides; Jacobian-vector products by difference quotient with sigma = 1/||v||_WRMS (CVLS default).
It is NOT CVODE: the BDF formulas use variable-coefficient divided differences on the stored solution history
instead of CVODE's fixed-leading-coefficient Nordsieck array, and the order/step heuristics are simpler, so
step sequences differ from CVODE's.  What it is for: running the RHS and the N_Vector together the way CVODE
does, comparing a GPU-driven run with an oracle-driven run of the identical integrator, and measuring
sim-days per wall-second.  All arithmetic on state vectors goes through `ops` (N_V* methods on device vectors).
"""
import math

import numpy as np


class BDFKrylov:
    def __init__(self, ops, new_vector, rhs, n, rtol=1e-4, atol=1e-4, max_step=10.0, init_step=1.0, min_step=1e-6,
                 max_order=2, maxl=5, n_global=None, linear_solver=None):
        """ops: N_V* table; new_vector(): allocate a device vector; rhs(t, y, ydot): CVRhsFn on device vectors.
        linear_solver: optional object with solve(t, gamma, y, fy, ewt, b, tol, x) -> (flag, nli, res), e.g. the
        device-resident DeviceSPGMR; default: the SPGMR below, written on the ops table."""
        self.linear_solver = linear_solver
        self.ops, self.new, self.rhs, self.n = ops, new_vector, rhs, n
        self.rtol, self.atol = rtol, atol
        self.hmax, self.h0, self.hmin = max_step, init_step, min_step
        self.qmax, self.maxl = max_order, maxl
        self.sqrtN = math.sqrt(n_global or n)
        v = self.new
        self.hist = [v() for _ in range(self.qmax + 2)]      # y_n, y_{n-1}, ... (ring, newest first)
        self.thist = []
        self.ewt, self.ypred, self.ycur, self.ftemp, self.b, self.delta, self.acor, self.tmp, self.tmp2 = (v() for _ in range(9))
        self.V = [v() for _ in range(maxl + 1)]
        self.q, self.h, self.t = 1, init_step, 0.0
        self.stats = dict(nst=0, nfe=0, nni=0, nli=0, netf=0, ncfn=0)
        self.nsuccess_at_q = 0

    # ---- helpers --------------------------------------------------------------------------------------
    def _f(self, t, y, out):
        self.rhs(t, y, out)
        self.stats["nfe"] += 1

    def _set_ewt(self, y):
        o = self.ops
        if hasattr(o, "EwtSet"):
            o.EwtSet(self.rtol, self.atol, y, self.ewt)
            return
        o.N_VAbs(y, self.ewt)
        o.N_VScale(self.rtol, self.ewt, self.ewt)
        o.N_VAddConst(self.ewt, self.atol, self.ewt)
        o.N_VInv(self.ewt, self.ewt)

    @staticmethod
    def _lagrange_weights(ts, t):
        """weights w_j with p(t) = sum w_j y_j for the polynomial through (ts_j, y_j)"""
        w = []
        for j, tj in enumerate(ts):
            num = den = 1.0
            for k, tk in enumerate(ts):
                if k != j:
                    num *= (t - tk); den *= (tj - tk)
            w.append(num / den)
        return w

    @staticmethod
    def _deriv_weights(ts, t):
        """weights a_j with p'(t) = sum a_j y_j (ts[0] == t is the new point)"""
        a = []
        m = len(ts)
        for j in range(m):
            s = 0.0
            for i in range(m):
                if i == j:
                    continue
                prod = 1.0 / (ts[j] - ts[i])
                for k in range(m):
                    if k != j and k != i:
                        prod *= (t - ts[k]) / (ts[j] - ts[k])
                s += prod
            a.append(s)
        return a

    def init(self, t0, y0):
        self.ops.N_VScale(1.0, y0, self.hist[0])
        self.thist = [t0]
        self.t, self.q, self.h = t0, 1, min(self.h0, self.hmax)
        self.nsuccess_at_q = 0
        self.f0 = self.new()
        self._f(t0, self.hist[0], self.f0)   # like CVODE's zn[1] = h f(t0, y0): the first predictor is Euler

    # ---- SPGMR on (I - gamma J) x = b, scaled by ewt on both sides -------------------------------------
    def _atimes(self, v, gamma, t, y, fy, out):
        """out = v - gamma * J v, J v by difference quotient (CVLS: sigma = 1/||v||_wrms)"""
        o = self.ops
        nrm = o.N_VWrmsNorm(v, self.ewt)
        if nrm == 0.0:
            o.N_VScale(1.0, v, out)
            return
        sig = 1.0 / nrm
        o.N_VLinearSum(sig, v, 1.0, y, self.tmp)
        self._f(t, self.tmp, self.tmp2)
        o.N_VLinearSum(1.0 / sig, self.tmp2, -1.0 / sig, fy, self.tmp2)   # J v
        o.N_VLinearSum(1.0, v, -gamma, self.tmp2, out)

    def _spgmr(self, b, x, gamma, t, y, fy, tol):
        """x <- approximate solution; returns (converged, iterations).  Scaled variables: xs = ewt*x."""
        o, V, maxl = self.ops, self.V, self.maxl
        o.N_VConst(0.0, x)
        o.N_VProd(self.ewt, b, V[0])                                      # scaled residual r0 = S b
        beta = math.sqrt(o.N_VDotProd(V[0], V[0]))
        if beta <= tol:
            return True, 0
        o.N_VScale(1.0 / beta, V[0], V[0])
        H = np.zeros((maxl + 1, maxl))
        g = np.zeros(maxl + 1); g[0] = beta
        cs, sn = np.zeros(maxl), np.zeros(maxl)
        k_used, conv = 0, False
        fused = hasattr(o, "DQPerturb")
        for k in range(maxl):
            self.stats["nli"] += 1
            if fused:
                # ||S^-1 v_k||_WRMS = ||v_k||_2 / sqrt(N) = 1/sqrt(N) exactly (v_k is normalised): CVLS' sigma
                # without a reduction; unscale+perturb and the whole scaled (I - gamma J) v in one launch each
                sig = self.sqrtN
                o.DQPerturb(sig, V[k], self.ewt, y, self.tmp)
                self._f(t, self.tmp, self.tmp2)
                o.DQCombine(sig, gamma, V[k], self.ewt, self.tmp2, fy, V[k + 1])
            else:
                o.N_VDiv(V[k], self.ewt, self.delta)                      # unscale: v = S^-1 v_k
                self._atimes(self.delta, gamma, t, y, fy, V[k + 1])
                o.N_VProd(self.ewt, V[k + 1], V[k + 1])                   # w = S A S^-1 v_k
            for i in range(k + 1):                                        # modified Gram-Schmidt
                H[i, k] = o.N_VDotProd(V[k + 1], V[i])
                o.N_VLinearSum(1.0, V[k + 1], -H[i, k], V[i], V[k + 1])
            H[k + 1, k] = math.sqrt(o.N_VDotProd(V[k + 1], V[k + 1]))
            if H[k + 1, k] != 0.0:
                o.N_VScale(1.0 / H[k + 1, k], V[k + 1], V[k + 1])
            for i in range(k):                                            # previous Givens rotations
                tmp = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = tmp
            den = math.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = (H[k, k] / den, H[k + 1, k] / den) if den else (1.0, 0.0)
            H[k, k] = cs[k] * H[k, k] + sn[k] * H[k + 1, k]
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            k_used = k + 1
            if abs(g[k + 1]) <= tol:
                conv = True
                break
        yk = np.zeros(k_used)
        for i in range(k_used - 1, -1, -1):
            yk[i] = (g[i] - H[i, i + 1:k_used] @ yk[i + 1:k_used]) / H[i, i]
        o.N_VLinearCombination(list(yk), V[:k_used], self.delta)          # scaled correction
        o.N_VDiv(self.delta, self.ewt, x)
        # CVODE accepts a reduced residual even if the tolerance was not met (SPGMR returns RES_REDUCED)
        return conv or abs(g[k_used]) < beta, k_used

    # ---- one step ---------------------------------------------------------------------------------------
    def step(self, tstop):
        o = self.ops
        ynew = self.hist[-1]                                              # slot that falls off the history
        while True:
            h = min(self.h, self.hmax, tstop - self.t)
            if h < self.hmin:
                h = min(self.hmin, tstop - self.t)
            q = min(self.q, len(self.thist))
            tn1 = self.t + h
            ts_old = self.thist[:q + 1]                                   # newest first
            # predictor: extrapolate the polynomial through the last (up to q+1) points
            if len(ts_old) == 1:
                o.N_VLinearSum(1.0, self.hist[0], h, self.f0, self.ypred)
            else:
                wp = self._lagrange_weights(ts_old, tn1)
                o.N_VLinearCombination(wp, self.hist[:len(ts_old)], self.ypred)
            self._set_ewt(self.hist[0])
            # BDF_q: p'(t_{n+1}) = f, p through y_{n+1} and the last q points
            a = self._deriv_weights([tn1] + self.thist[:q], tn1)
            gamma = 1.0 / a[0]
            # rhs constant: psi = -(sum_{j>=1} a_j y_{n+1-j}) / a0  =>  G(y) = y - gamma f(y) - psi
            o.N_VLinearCombination([-aj * gamma for aj in a[1:]], self.hist[:q], self.b)
            o.N_VScale(1.0, self.ypred, self.ycur)
            o.N_VConst(0.0, self.acor)
            conv, crate, delp = False, 1.0, 0.0
            for mnewt in range(3):
                self.stats["nni"] += 1
                self._f(tn1, self.ycur, self.ftemp)
                # residual of the Newton system: r = gamma f + psi - y
                rhsvec = self.V[self.maxl]                                # borrowed: free until _spgmr normalises V[0]
                if hasattr(o, "NewtonResid"):
                    o.NewtonResid(gamma, self.ftemp, self.b, self.ycur, rhsvec)
                else:
                    o.N_VLinearSum(gamma, self.ftemp, 1.0, self.b, self.tmp)
                    o.N_VLinearSum(1.0, self.tmp, -1.0, self.ycur, rhsvec)
                x = self.hist[-1]                                         # scratch: the slot about to be overwritten
                if self.linear_solver is not None:
                    flag, nli, _ = self.linear_solver.solve(tn1, gamma, self.ycur, self.ftemp, self.ewt, rhsvec,
                                                            0.05 * 0.1 * self.sqrtN, x)
                    ok = flag <= 1
                    self.stats["nli"] += nli
                    self.stats["nfe"] += nli
                else:
                    ok, _ = self._spgmr(rhsvec, x, gamma, tn1, self.ycur, self.ftemp, 0.05 * 0.1 * self.sqrtN)
                if hasattr(o, "NewtonUpdate"):
                    dele = o.NewtonUpdate(x, self.ewt, self.ycur, self.acor)
                else:
                    dele = o.N_VWrmsNorm(x, self.ewt)
                    o.N_VLinearSum(1.0, self.ycur, 1.0, x, self.ycur)
                    o.N_VLinearSum(1.0, self.acor, 1.0, x, self.acor)
                if mnewt > 0:
                    crate = max(0.3 * crate, dele / delp) if delp > 0 else crate
                dcon = dele * min(1.0, crate) / 0.1
                if dcon <= 1.0 and ok:
                    conv = True
                    break
                if mnewt > 0 and dele > 2.0 * delp:
                    break
                delp = dele
            if not conv:
                self.stats["ncfn"] += 1
                self.h = max(h * 0.25, self.hmin)
                if h <= self.hmin * 1.0000001:
                    raise RuntimeError("Newton iteration failed at the minimum step")
                continue
            # local error estimate: (h / (t_{n+1} - t_{n-q})) * ||y - y_pred||  (variable-step BDF_q)
            span = tn1 - ts_old[min(q, len(ts_old) - 1)]
            o.N_VLinearSum(1.0, self.ycur, -1.0, self.ypred, self.tmp)
            err = (h / span) * o.N_VWrmsNorm(self.tmp, self.ewt) if len(ts_old) > q else o.N_VWrmsNorm(self.tmp, self.ewt) * 0.5
            if err > 1.0:
                self.stats["netf"] += 1
                fac = max(0.2, 0.9 * err ** (-1.0 / (q + 1)))
                self.h = max(h * fac, self.hmin)
                if h <= self.hmin * 1.0000001:
                    raise RuntimeError("error test failed at the minimum step")
                self.nsuccess_at_q = 0
                continue
            # accept
            o.N_VScale(1.0, self.ycur, ynew)
            self.hist = [ynew] + self.hist[:-1]
            self.thist = [tn1] + self.thist[:self.qmax + 1]
            self.t = tn1
            self.stats["nst"] += 1
            self.nsuccess_at_q += 1
            fac = min(2.0, max(0.2, 0.9 * max(err, 1e-10) ** (-1.0 / (q + 1))))
            if fac < 1.0 or fac > 1.2:
                self.h = h * fac
            else:
                self.h = h
            if self.q < self.qmax and self.nsuccess_at_q > self.q + 1 and len(self.thist) > self.q + 1:
                self.q += 1
                self.nsuccess_at_q = 0
            return self.t

    def advance(self, tout):
        """CVode(mem, tout, ..., CV_NORMAL) with a stop time at tout; returns the state vector at tout"""
        while self.t < tout - 1e-10:
            self.step(tout)
        return self.hist[0]

    def reset_history(self):
        """cold restart of the multistep history (what a forcing discontinuity calls for)"""
        self.thist = self.thist[:1]
        self.q, self.nsuccess_at_q = 1, 0
        self._f(self.t, self.hist[0], self.f0)
