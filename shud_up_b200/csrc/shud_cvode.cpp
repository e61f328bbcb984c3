// shud_cvode.cpp - the integrator side of the boundary (include/shud_cvode.h): CVODE's variable-order, variable-step
// BDF in fixed-leading-coefficient Nordsieck form with Newton + matrix-free SPGMR, restated for exactly the
// configuration the reference runs (src/Equations/cvode_config.cpp:162-193, src/Model/shud.cpp:89-133) and written
// ONLY against the generic N_Vector operations (N_VLinearSum, N_VWrmsNorm, ... -> v->ops), so the same object code
// drives the device vector + shud_b200_f and a host vector + the CPU oracle.
//
// SUNDIALS is third-party and absent from the reference tree and from this image (configure:17-21 pins
// cvode-6.0.0).  What follows restates the published algorithm of CVODE 6 (Hindmarsh et al., "SUNDIALS: Suite of
// Nonlinear and Differential/Algebraic Equation Solvers", ACM TOMS 31(3), 2005; Brown, Byrne, Hindmarsh, "VODE",
// SIAM J. Sci. Stat. Comput. 10, 1989; the CVODE 6 user guide, chapter "Mathematical considerations") with the
// library's default constants.  Function names in the comments (cvStep, cvSetBDF, ...) are CVODE's, so a reader
// can lay the two side by side.  "parity unpinned": no reference test pins step sequences.
//
// Plain host code (no CUDA): compiled into libshud_b200.so and into the CPU-only checker library.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "shud_cvode.h"

namespace {

// ---- CVODE's constants (cvode_impl.h / cvode.c) ----
constexpr int Q_MAX = 5, L_MAX = Q_MAX + 1, NUM_TESTS = 5;
constexpr double FUZZ_FACTOR = 100.0, HLB_FACTOR = 100.0, HUB_FACTOR = 0.1, H_BIAS = 0.5;
constexpr int MAX_HIN_ITERS = 4;
constexpr double ETAMX1 = 10000.0, ETAMX2 = 10.0, ETAMX3 = 10.0, ETAMXF = 0.2, ETAMIN = 0.1, ETACF = 0.25;
constexpr double ADDON = 1e-6, BIAS1 = 6.0, BIAS2 = 6.0, BIAS3 = 10.0, THRESH = 1.5, ONEPSM = 1.000001;
constexpr int SMALL_NST = 10, MXNCF = 10, MXNEF = 7, MXNEF1 = 3, SMALL_NEF = 2, LONG_WAIT = 10;
constexpr int NLS_MAXCOR = 3;
constexpr double CRDOWN = 0.3, RDIV = 2.0, NLSCOEF = 0.1;
constexpr double EPLIFAC = 0.05;   // CVLS_EPLIN
constexpr int MAX_DQITERS = 3;
constexpr int SPGMR_MAXL_DEFAULT = 5;
constexpr double GS_FACTOR = 1000.0;

// internal step / solver flags
enum { DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, TRY_AGAIN = 5, FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8 };
enum { NLS_SUCCESS = 0, NLS_CONTINUE = 901, NLS_CONV_RECVR = 902 };
enum { LS_SUCCESS = 0, LS_RES_REDUCED = 801, LS_CONV_FAIL = 802, LS_QRFACT_FAIL = 806, LS_ATIMES_FAIL = -1, LS_QRSOL_FAIL = -2 };

}  // namespace

struct shud_cv {
    shud_cv_rhs_fn f;
    void *user_data;
    // problem / tolerances
    double reltol, abstol;
    int tol_set;
    // Nordsieck history and work vectors (cloned from y0)
    N_Vector zn[L_MAX + 1], ewt, y, acor, tempv, ftemp, vtemp1;
    // step data
    int q, qprime, next_q, qwait, L, qmax, qu;
    double hin, h, hprime, next_h, eta, hscale, tn, tretlast, hu, h0u;
    double tau[L_MAX + 1], tq[NUM_TESTS + 1], l[L_MAX + 1];
    double rl1, gamma, gammap, gamrat, crate, delp, acnrm, nlscoef;
    int acnrmcur, mnewt;
    double hmin, hmax_inv, etamax, saved_tq5, tstop;
    int tstopset, indx_acor;
    long mxstep, nst, nfe, ncfn, netf, nni, nscon;
    // linear solver (SPGMR, matrix-free)
    int maxl;
    N_Vector V[SPGMR_MAXL_DEFAULT + 2], xcor, ls_x, ls_vtemp;
    double Hes[SPGMR_MAXL_DEFAULT + 2][SPGMR_MAXL_DEFAULT + 1], givens[2 * (SPGMR_MAXL_DEFAULT + 1)], yg[SPGMR_MAXL_DEFAULT + 2];
    double nrmfac;
    long nli, ncfl, nfeLS, nps;
    shud_cv_fused fused;
    int have_fused;
    int nls_primed;  // the fused predictor has left acor = 0 and y = zn[0] + acor
    // fused cvCompleteStep: the weights and the tolsf norm of the NEXT step are formed with the update of zn (the step's
    // own weights stay in ewt for cvPrepareNextStep); the top of CVode()'s loop swaps them in
    N_Vector ewt_next;
    int ewt_next_valid;
    double nrm_next;
    N_Vector yout_hint;  // CV_ONE_STEP: the caller's vector, filled by the same pass
    int yout_done;
    double uround;
};

namespace {

inline double rabs(double x) { return fabs(x); }
inline double rmax(double a, double b) { return a > b ? a : b; }
inline double rmin(double a, double b) { return a < b ? a : b; }
double rpowerI(double base, int e) {
    double p = 1.0;
    const int n = e < 0 ? -e : e;
    for (int i = 0; i < n; i++) p *= base;
    return e < 0 ? 1.0 / p : p;
}
inline double rpowerR(double base, double e) { return base <= 0.0 ? 0.0 : pow(base, e); }

// cvEwtSetSS: ewt = 1 / (reltol |y| + abstol)
int ewt_set(shud_cv *cv, N_Vector ycur, N_Vector weight) {
    if (cv->have_fused && cv->fused.ewt_set && cv->abstol > 0.0)
        return cv->fused.ewt_set(cv->fused.ctx, cv->reltol, cv->abstol, ycur, weight);
    N_VAbs(ycur, cv->tempv);
    N_VScale(cv->reltol, cv->tempv, cv->tempv);
    N_VAddConst(cv->tempv, cv->abstol, cv->tempv);
    if (cv->abstol == 0.0) {
        if (N_VMin(cv->tempv) <= 0.0) return -1;
    }
    N_VInv(cv->tempv, weight);
    return 0;
}

// ---------------------------------------------------------------- linear solver: CVLS + SPGMR ----
// cvLsDQJtimes: Jv = (f(t, y + sig v) - fy) / sig, sig = 1 / ||v||_WRMS
int dq_jtimes(shud_cv *cv, N_Vector v, N_Vector Jv, double t, N_Vector y, N_Vector fy, N_Vector work) {
    double sig = 1.0 / N_VWrmsNorm(v, cv->ewt);
    int retval = 0;
    for (int iter = 0; iter < MAX_DQITERS; iter++) {
        N_VLinearSum(sig, v, 1.0, y, work);
        retval = cv->f(t, work, Jv, cv->user_data);
        cv->nfeLS++;
        if (retval == 0) break;
        if (retval < 0) return -1;
        sig *= 0.25;
    }
    if (retval > 0) return 1;
    const double siginv = 1.0 / sig;
    N_VLinearSum(siginv, Jv, -siginv, fy, Jv);
    return 0;
}

// cvLsATimes: z = (I - gamma J) v
int atimes(shud_cv *cv, N_Vector v, N_Vector z, N_Vector ycur, N_Vector fcur) {
    const int r = dq_jtimes(cv, v, z, cv->tn, ycur, fcur, cv->vtemp1);
    if (r != 0) return r;
    N_VLinearSum(1.0, v, -cv->gamma, z, z);
    return 0;
}

// Givens QR of the Hessenberg matrix (SUNQRfact, job 0 on the first column, update afterwards) and its solve
int qr_fact(int n, double h[][SPGMR_MAXL_DEFAULT + 1], double *q, int job) {
    double c, s, temp1, temp2, temp3;
    int code = 0;
    if (job == 0) {
        for (int k = 0; k < n; k++) {
            for (int j = 0; j < k - 1; j++) {
                const int i = 2 * j;
                temp1 = h[j][k]; temp2 = h[j + 1][k];
                c = q[i]; s = q[i + 1];
                h[j][k] = c * temp1 - s * temp2;
                h[j + 1][k] = s * temp1 + c * temp2;
            }
            const int q_ptr = 2 * k;
            temp1 = h[k][k]; temp2 = h[k + 1][k];
            if (temp2 == 0.0) { c = 1.0; s = 0.0; }
            else if (rabs(temp2) >= rabs(temp1)) { temp3 = temp1 / temp2; s = -1.0 / sqrt(1.0 + temp3 * temp3); c = -s * temp3; }
            else { temp3 = temp2 / temp1; c = 1.0 / sqrt(1.0 + temp3 * temp3); s = -c * temp3; }
            q[q_ptr] = c; q[q_ptr + 1] = s;
            if ((h[k][k] = c * temp1 - s * temp2) == 0.0) code = k + 1;
        }
    } else {
        const int n_minus_1 = n - 1;
        for (int k = 0; k < n_minus_1; k++) {
            const int i = 2 * k;
            temp1 = h[k][n_minus_1]; temp2 = h[k + 1][n_minus_1];
            c = q[i]; s = q[i + 1];
            h[k][n_minus_1] = c * temp1 - s * temp2;
            h[k + 1][n_minus_1] = s * temp1 + c * temp2;
        }
        temp1 = h[n_minus_1][n_minus_1]; temp2 = h[n][n_minus_1];
        if (temp2 == 0.0) { c = 1.0; s = 0.0; }
        else if (rabs(temp2) >= rabs(temp1)) { temp3 = temp1 / temp2; s = -1.0 / sqrt(1.0 + temp3 * temp3); c = -s * temp3; }
        else { temp3 = temp2 / temp1; c = 1.0 / sqrt(1.0 + temp3 * temp3); s = -c * temp3; }
        const int q_ptr = 2 * n_minus_1;
        q[q_ptr] = c; q[q_ptr + 1] = s;
        if ((h[n_minus_1][n_minus_1] = c * temp1 - s * temp2) == 0.0) code = n;
    }
    return code;
}
int qr_sol(int n, double h[][SPGMR_MAXL_DEFAULT + 1], double *q, double *b) {
    for (int k = 0; k < n; k++) {
        const int q_ptr = 2 * k;
        const double c = q[q_ptr], s = q[q_ptr + 1], temp1 = b[k], temp2 = b[k + 1];
        b[k] = c * temp1 - s * temp2;
        b[k + 1] = s * temp1 + c * temp2;
    }
    for (int k = n - 1; k >= 0; k--) {
        if (h[k][k] == 0.0) return k + 1;
        b[k] /= h[k][k];
        for (int i = 0; i < k; i++) b[i] -= b[k] * h[i][k];
    }
    return 0;
}

// SUNModifiedGS: orthogonalise v[k] against v[max(k-p,0)..k-1], with the re-orthogonalisation test
void modified_gs(N_Vector *v, double h[][SPGMR_MAXL_DEFAULT + 1], int k, int p, double *new_vk_norm) {
    const int k_minus_1 = k - 1, i0 = k - p > 0 ? k - p : 0;
    const double vk_norm = sqrt(N_VDotProd(v[k], v[k]));
    for (int i = i0; i < k; i++) {
        h[i][k_minus_1] = N_VDotProd(v[i], v[k]);
        N_VLinearSum(1.0, v[k], -h[i][k_minus_1], v[i], v[k]);
    }
    *new_vk_norm = sqrt(N_VDotProd(v[k], v[k]));
    double temp = GS_FACTOR * vk_norm;
    if ((temp + (*new_vk_norm)) != temp) return;
    double new_norm_2 = 0.0;
    for (int i = i0; i < k; i++) {
        const double new_product = N_VDotProd(v[i], v[k]);
        temp = GS_FACTOR * h[i][k_minus_1];
        if ((temp + new_product) == temp) continue;
        h[i][k_minus_1] += new_product;
        N_VLinearSum(1.0, v[k], -new_product, v[i], v[k]);
        new_norm_2 += new_product * new_product;
    }
    if (new_norm_2 != 0.0) {
        const double new_product = (*new_vk_norm) * (*new_vk_norm) - new_norm_2;
        *new_vk_norm = new_product > 0.0 ? sqrt(new_product) : 0.0;
    }
}

// SUNLinSolSolve_SPGMR for this configuration: no preconditioner, scaling s1 = s2 = ewt, zero initial guess,
// modified Gram-Schmidt, no restarts.  x receives the solution.
int spgmr_solve(shud_cv *cv, N_Vector x, N_Vector b, double delta, N_Vector ycur, N_Vector fcur, int *nli_out) {
    const int l_max = cv->maxl;
    N_Vector *V = cv->V, xcor = cv->xcor, vtemp = cv->ls_vtemp, s = cv->ewt;
    int krydim = 0, l_plus_1 = 0;
    bool converged = false;
    *nli_out = 0;
    for (int i = 0; i <= l_max; i++)
        for (int j = 0; j < l_max; j++) cv->Hes[i][j] = 0.0;
    // r_0 = b (zero guess), scaled: V[0] = s1 r_0
    N_VProd(s, b, V[0]);
    double r_norm = sqrt(N_VDotProd(V[0], V[0]));
    const double beta = r_norm;
    double rho = beta;
    if (r_norm <= delta) { N_VConst(0.0, x); return LS_SUCCESS; }
    double rotation_product = 1.0;
    N_VScale(1.0 / r_norm, V[0], V[0]);
    N_VConst(0.0, xcor);
    int l;
    for (l = 0; l < l_max; l++) {
        (*nli_out)++;
        krydim = l_plus_1 = l + 1;
        // A-tilde V[l] = s1 A s2^{-1} V[l]
        N_VDiv(V[l], s, vtemp);
        const int ier = atimes(cv, vtemp, V[l_plus_1], ycur, fcur);
        if (ier != 0) return ier < 0 ? LS_ATIMES_FAIL : LS_CONV_FAIL;
        N_VProd(s, V[l_plus_1], V[l_plus_1]);
        modified_gs(V, cv->Hes, l_plus_1, l_max, &cv->Hes[l_plus_1][l]);
        if (qr_fact(krydim, cv->Hes, cv->givens, l) != 0) return LS_QRFACT_FAIL;
        rotation_product *= cv->givens[2 * l + 1];
        rho = rabs(rotation_product * r_norm);
        if (rho <= delta) { converged = true; break; }
        N_VScale(1.0 / cv->Hes[l_plus_1][l], V[l_plus_1], V[l_plus_1]);
    }
    // least-squares solution of the small problem, correction xcor = V y
    cv->yg[0] = r_norm;
    for (int i = 1; i <= krydim; i++) cv->yg[i] = 0.0;
    if (qr_sol(krydim, cv->Hes, cv->givens, cv->yg) != 0) return LS_QRSOL_FAIL;
    {
        double cvals[SPGMR_MAXL_DEFAULT + 2];
        N_Vector Xv[SPGMR_MAXL_DEFAULT + 2];
        cvals[0] = 1.0; Xv[0] = xcor;
        for (int k = 0; k < krydim; k++) { cvals[k + 1] = cv->yg[k]; Xv[k + 1] = V[k]; }
        if (N_VLinearCombination(krydim + 1, cvals, Xv, xcor) != 0) return LS_ATIMES_FAIL;
    }
    if (converged || rho < beta) {
        N_VDiv(xcor, s, x);  // x = s2^{-1} xcor (zero guess)
        return converged ? LS_SUCCESS : LS_RES_REDUCED;
    }
    return LS_CONV_FAIL;
}

// cvLsSolve: b holds the right-hand side on entry and the solution on return.  0 ok, > 0 recoverable, < 0 fatal.
int ls_solve(shud_cv *cv, N_Vector b, N_Vector ynow, N_Vector fnow) {
    const int curiter = cv->mnewt;
    const double deltar = EPLIFAC * cv->tq[4];
    if (cv->have_fused && cv->fused.lsolve) {
        // the fused solver does the norm test on its own scaled residual: ||b||_WRMS <= deltar <=> ||s b||_2 <= delta
        int nli = 0, nfe = 0;
        const double delta = deltar * cv->nrmfac;
        const int r = cv->fused.lsolve(cv->fused.ctx, cv->tn, cv->gamma, ynow, fnow, cv->ewt, b, delta, cv->ls_x, &nli, &nfe);
        cv->nli += nli; cv->nfeLS += nfe;
        if (r < 0) return -1;
        if (r == 3) {  // right-hand side already below the tolerance: x = 0 (later iterations) or x = b (first)
            if (curiter > 0) N_VConst(0.0, b);
            return 0;
        }
        N_VScale(1.0, cv->ls_x, b);
        if (r != 0) cv->ncfl++;
        if (r == 0) return 0;
        if (r == 1) return curiter == 0 ? 0 : 1;
        return 1;
    }
    const double bnorm = N_VWrmsNorm(b, cv->ewt);
    if (bnorm <= deltar) {
        if (curiter > 0) N_VConst(0.0, b);
        return 0;
    }
    const double delta = deltar * cv->nrmfac;
    int nli = 0;
    const int retval = spgmr_solve(cv, cv->ls_x, b, delta, ynow, fnow, &nli);
    N_VScale(1.0, cv->ls_x, b);
    cv->nli += nli;
    if (retval != LS_SUCCESS) cv->ncfl++;
    switch (retval) {
        case LS_SUCCESS: return 0;
        case LS_RES_REDUCED: return curiter == 0 ? 0 : 1;
        case LS_CONV_FAIL: case LS_QRFACT_FAIL: return 1;
        default: return -1;
    }
}

// ---------------------------------------------------------------- nonlinear solver: Newton ----
// res = rl1 zn[1] + ycor - gamma ftemp
int residual_of_ftemp(shud_cv *cv, N_Vector ycor, N_Vector res) {
    if (cv->have_fused && cv->fused.nls_residual)
        return cv->fused.nls_residual(cv->fused.ctx, cv->rl1, cv->zn[1], ycor, cv->gamma, cv->ftemp, res) ? SHUD_CV_RHSFUNC_FAIL : 0;
    N_VLinearSum(cv->rl1, cv->zn[1], 1.0, ycor, res);
    N_VLinearSum(-cv->gamma, cv->ftemp, 1.0, res, res);
    return 0;
}

// cvNlsResidual: res = rl1 zn[1] + ycor - gamma f(tn, zn[0] + ycor); y_current: cv->y already holds zn[0] + ycor
int nls_residual(shud_cv *cv, N_Vector ycor, N_Vector res, bool y_current = false) {
    if (!y_current) N_VLinearSum(1.0, cv->zn[0], 1.0, ycor, cv->y);
    const int retval = cv->f(cv->tn, cv->y, cv->ftemp, cv->user_data);
    cv->nfe++;
    if (retval < 0) return SHUD_CV_RHSFUNC_FAIL;
    if (retval > 0) return 1;
    return residual_of_ftemp(cv, ycor, res);
}

// cvNlsConvTest (del = ||delta||_WRMS)
int nls_conv_test(shud_cv *cv, N_Vector ycor, double del, double tol) {
    const int m = cv->mnewt;
    if (m > 0) cv->crate = rmax(CRDOWN * cv->crate, del / cv->delp);
    const double dcon = del * rmin(1.0, cv->crate) / tol;
    if (dcon <= 1.0) {
        cv->acnrm = (m == 0) ? del : N_VWrmsNorm(ycor, cv->ewt);
        cv->acnrmcur = 1;
        return NLS_SUCCESS;
    }
    if (m >= 1 && del > RDIV * cv->delp) return NLS_CONV_RECVR;
    cv->delp = del;
    return NLS_CONTINUE;
}

// cvNls + SUNNonlinSolSolve_Newton.  With a matrix-free linear solver and no preconditioner CVLS leaves no setup
// routine (cvLsInitialize), so crate restarts at 1 on every call and there is no "retry with a fresh Jacobian".
int nls(shud_cv *cv) {
    cv->crate = 1.0;
    const bool primed = cv->nls_primed != 0;  // the fused predictor has already set acor = 0, y = zn[0] + acor
    cv->nls_primed = 0;
    if (!primed) N_VConst(0.0, cv->acor);
    cv->acnrmcur = 0;
    N_Vector delta = cv->tempv;
    cv->mnewt = 0;
    if (cv->have_fused && cv->fused.newton_step) {
        // Device route: the residual, the linear solve and the update of one iteration are a single hook; y is
        // always zn[0] + acor when f is evaluated, as in nls_residual.  Same arithmetic as the route below.
        if (!primed) N_VLinearSum(1.0, cv->zn[0], 1.0, cv->acor, cv->y);
        for (;;) {
            int retval = cv->f(cv->tn, cv->y, cv->ftemp, cv->user_data);
            cv->nfe++;
            if (retval < 0) return SHUD_CV_RHSFUNC_FAIL;
            if (retval > 0) return NLS_CONV_RECVR;
            cv->nni++;
            double del = 0.0;
            int nli = 0, nfe = 0;
            const int r = cv->fused.newton_step(cv->fused.ctx, cv->tn, cv->gamma, cv->rl1, cv->zn[0], cv->zn[1], cv->acor, cv->y,
                                                cv->ftemp, cv->ewt, EPLIFAC * cv->tq[4] * cv->nrmfac, &del, &nli, &nfe);
            cv->nli += nli; cv->nfeLS += nfe;
            if (r < 0) return SHUD_CV_LSOLVE_FAIL;
            if (r == 3) {
                // right-hand side already below the tolerance (cvLsSolve's norm test): delta = b on the first
                // iteration, 0 later - rare, through the vector operations
                if (residual_of_ftemp(cv, cv->acor, delta)) return SHUD_CV_RHSFUNC_FAIL;
                N_VScale(-1.0, delta, delta);
                if (cv->mnewt > 0) N_VConst(0.0, delta);
                N_VLinearSum(1.0, cv->acor, 1.0, delta, cv->acor);
                N_VLinearSum(1.0, cv->zn[0], 1.0, cv->acor, cv->y);
                del = N_VWrmsNorm(delta, cv->ewt);
            } else {
                if (r != 0) cv->ncfl++;
                if (r == 2 || (r == 1 && cv->mnewt > 0)) return NLS_CONV_RECVR;
            }
            retval = nls_conv_test(cv, cv->acor, del, cv->tq[4]);
            if (retval == NLS_SUCCESS) break;
            if (retval != NLS_CONTINUE) return retval;
            cv->mnewt++;
            if (cv->mnewt >= NLS_MAXCOR) return NLS_CONV_RECVR;
        }
        if (!cv->acnrmcur) cv->acnrm = N_VWrmsNorm(cv->acor, cv->ewt);
        return 0;
    }
    int retval = nls_residual(cv, cv->acor, delta, primed);
    if (retval != 0) return retval < 0 ? retval : NLS_CONV_RECVR;
    for (;;) {
        cv->nni++;
        N_VScale(-1.0, delta, delta);
        retval = ls_solve(cv, delta, cv->y, cv->ftemp);
        if (retval < 0) return SHUD_CV_LSOLVE_FAIL;
        if (retval > 0) return NLS_CONV_RECVR;
        N_VLinearSum(1.0, cv->acor, 1.0, delta, cv->acor);
        retval = nls_conv_test(cv, cv->acor, N_VWrmsNorm(delta, cv->ewt), cv->tq[4]);
        if (retval == NLS_SUCCESS) break;
        if (retval != NLS_CONTINUE) return retval;
        cv->mnewt++;
        if (cv->mnewt >= NLS_MAXCOR) return NLS_CONV_RECVR;
        retval = nls_residual(cv, cv->acor, delta);
        if (retval != 0) return retval < 0 ? retval : NLS_CONV_RECVR;
    }
    N_VLinearSum(1.0, cv->zn[0], 1.0, cv->acor, cv->y);
    if (!cv->acnrmcur) cv->acnrm = N_VWrmsNorm(cv->acor, cv->ewt);
    return 0;
}

// ---------------------------------------------------------------- the step ----
void rescale(shud_cv *cv) {  // cvRescale
    double factor = cv->eta;
    for (int j = 1; j <= cv->q; j++) {
        N_VScale(factor, cv->zn[j], cv->zn[j]);
        factor *= cv->eta;
    }
    cv->h = cv->hscale * cv->eta;
    cv->next_h = cv->h;
    cv->hscale = cv->h;
    cv->nscon = 0;
}

void increase_bdf(shud_cv *cv) {  // cvIncreaseBDF
    for (int i = 0; i <= cv->qmax; i++) cv->l[i] = 0.0;
    double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, hsum = cv->hscale;
    cv->l[2] = 1.0;
    if (cv->q > 1) {
        for (int j = 1; j < cv->q; j++) {
            hsum += cv->tau[j + 1];
            const double xi = hsum / cv->hscale;
            prod *= xi;
            alpha0 -= 1.0 / (j + 1);
            alpha1 += 1.0 / xi;
            for (int i = j + 2; i >= 2; i--) cv->l[i] = cv->l[i] * xiold + cv->l[i - 1];
            xiold = xi;
        }
    }
    const double A1 = (-alpha0 - alpha1) / prod;
    N_VScale(A1, cv->zn[cv->indx_acor], cv->zn[cv->L]);
    for (int j = 2; j <= cv->q; j++) N_VLinearSum(cv->l[j], cv->zn[cv->L], 1.0, cv->zn[j], cv->zn[j]);
}

void decrease_bdf(shud_cv *cv) {  // cvDecreaseBDF
    for (int i = 0; i <= cv->qmax; i++) cv->l[i] = 0.0;
    cv->l[2] = 1.0;
    double hsum = 0.0;
    for (int j = 1; j <= cv->q - 2; j++) {
        hsum += cv->tau[j];
        const double xi = hsum / cv->hscale;
        for (int i = j + 2; i >= 2; i--) cv->l[i] = cv->l[i] * xi + cv->l[i - 1];
    }
    for (int j = 2; j < cv->q; j++) N_VLinearSum(-cv->l[j], cv->zn[cv->q], 1.0, cv->zn[j], cv->zn[j]);
}

void adjust_order(shud_cv *cv, int deltaq) {  // cvAdjustOrder (BDF)
    if (cv->q == 2 && deltaq != 1) return;
    if (deltaq == 1) increase_bdf(cv);
    else if (deltaq == -1) decrease_bdf(cv);
}

void adjust_params(shud_cv *cv) {  // cvAdjustParams
    if (cv->qprime != cv->q) {
        adjust_order(cv, cv->qprime - cv->q);
        cv->q = cv->qprime;
        cv->L = cv->q + 1;
        cv->qwait = cv->L;
    }
    rescale(cv);
}

void predict(shud_cv *cv) {  // cvPredict
    cv->tn += cv->h;
    if (cv->tstopset) {
        if ((cv->tn - cv->tstop) * cv->h > 0.0) cv->tn = cv->tstop;
    }
    if (cv->have_fused && cv->fused.predict) {
        // one pass over the Nordsieck array; it also leaves acor = 0 and y = zn[0] + acor for the Newton iteration
        if (cv->fused.predict(cv->fused.ctx, cv->q, 1.0, cv->zn, cv->y, cv->acor) == 0) { cv->nls_primed = 1; return; }
    }
    for (int k = 1; k <= cv->q; k++)
        for (int j = cv->q; j >= k; j--) N_VLinearSum(1.0, cv->zn[j - 1], 1.0, cv->zn[j], cv->zn[j - 1]);
}

void restore(shud_cv *cv, double saved_t) {  // cvRestore
    cv->tn = saved_t;
    if (cv->have_fused && cv->fused.predict && cv->fused.predict(cv->fused.ctx, cv->q, -1.0, cv->zn, nullptr, nullptr) == 0) return;
    for (int k = 1; k <= cv->q; k++)
        for (int j = cv->q; j >= k; j--) N_VLinearSum(1.0, cv->zn[j - 1], -1.0, cv->zn[j], cv->zn[j - 1]);
}

void set_tq_bdf(shud_cv *cv, double hsum, double alpha0, double alpha0_hat, double xi_inv, double xistar_inv) {  // cvSetTqBDF
    const int q = cv->q;
    const double A1 = 1.0 - alpha0_hat + alpha0;
    const double A2 = 1.0 + q * A1;
    cv->tq[2] = rabs(A1 / (alpha0 * A2));
    cv->tq[5] = rabs(A2 * xistar_inv / (cv->l[q] * xi_inv));
    if (cv->qwait == 1) {
        if (q > 1) {
            const double C = xistar_inv / cv->l[q];
            const double A3 = alpha0 + 1.0 / q;
            const double A4 = alpha0_hat + xi_inv;
            const double Cpinv = (1.0 - A4 + A3) / A3;
            cv->tq[1] = rabs(C * Cpinv);
        } else {
            cv->tq[1] = 1.0;
        }
        hsum += cv->tau[q];
        xi_inv = cv->h / hsum;
        const double A5 = alpha0 - (1.0 / (q + 1));
        const double A6 = alpha0_hat - xi_inv;
        const double Cppinv = (1.0 - A6 + A5) / A2;
        cv->tq[3] = rabs(Cppinv / (xi_inv * (q + 2) * A5));
    }
    cv->tq[4] = cv->nlscoef / cv->tq[2];
}

void set_bdf(shud_cv *cv) {  // cvSetBDF
    const int q = cv->q;
    double xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0, hsum = cv->h;
    cv->l[0] = cv->l[1] = 1.0;
    for (int i = 2; i <= q; i++) cv->l[i] = 0.0;
    if (q > 1) {
        for (int j = 2; j < q; j++) {
            hsum += cv->tau[j - 1];
            xi_inv = cv->h / hsum;
            alpha0 -= 1.0 / j;
            for (int i = j; i >= 1; i--) cv->l[i] += cv->l[i - 1] * xi_inv;
        }
        alpha0 -= 1.0 / q;
        xistar_inv = -cv->l[1] - alpha0;
        hsum += cv->tau[q - 1];
        xi_inv = cv->h / hsum;
        alpha0_hat = -cv->l[1] - xi_inv;
        for (int i = q; i >= 1; i--) cv->l[i] += cv->l[i - 1] * xistar_inv;
    }
    set_tq_bdf(cv, hsum, alpha0, alpha0_hat, xi_inv, xistar_inv);
}

void set_coeffs(shud_cv *cv) {  // cvSet
    set_bdf(cv);
    cv->rl1 = 1.0 / cv->l[1];
    cv->gamma = cv->h * cv->rl1;
    if (cv->nst == 0) cv->gammap = cv->gamma;
    cv->gamrat = (cv->nst > 0) ? cv->gamma / cv->gammap : 1.0;
}

int handle_nflag(shud_cv *cv, int *nflagPtr, double saved_t, int *ncfPtr) {  // cvHandleNFlag
    const int nflag = *nflagPtr;
    if (nflag == 0) return DO_ERROR_TEST;
    cv->ncfn++;
    restore(cv, saved_t);
    if (nflag < 0) return nflag;  // SHUD_CV_LSOLVE_FAIL, SHUD_CV_RHSFUNC_FAIL
    (*ncfPtr)++;
    cv->etamax = 1.0;
    if (rabs(cv->h) <= cv->hmin * ONEPSM || *ncfPtr == MXNCF) return SHUD_CV_CONV_FAILURE;
    cv->eta = rmax(ETACF, cv->hmin / rabs(cv->h));
    *nflagPtr = PREV_CONV_FAIL;
    rescale(cv);
    return PREDICT_AGAIN;
}

int do_error_test(shud_cv *cv, int *nflagPtr, double saved_t, int *nefPtr, double *dsmPtr) {  // cvDoErrorTest
    const double dsm = cv->acnrm * cv->tq[2];
    *dsmPtr = dsm;
    if (dsm <= 1.0) return 0;
    (*nefPtr)++;
    cv->netf++;
    *nflagPtr = PREV_ERR_FAIL;
    restore(cv, saved_t);
    if (rabs(cv->h) <= cv->hmin * ONEPSM || *nefPtr == MXNEF) return SHUD_CV_ERR_FAILURE;
    cv->etamax = 1.0;
    if (*nefPtr <= MXNEF1) {
        cv->eta = 1.0 / (rpowerR(BIAS2 * dsm, 1.0 / cv->L) + ADDON);
        cv->eta = rmax(ETAMIN, rmax(cv->eta, cv->hmin / rabs(cv->h)));
        if (*nefPtr >= SMALL_NEF) cv->eta = rmin(cv->eta, ETAMXF);
        rescale(cv);
        return TRY_AGAIN;
    }
    if (cv->q > 1) {  // after MXNEF1 failures: force an order reduction
        cv->eta = rmax(ETAMIN, cv->hmin / rabs(cv->h));
        adjust_order(cv, -1);
        cv->L = cv->q;
        cv->q--;
        cv->qwait = cv->L;
        rescale(cv);
        return TRY_AGAIN;
    }
    // already at order 1: reload the history from scratch
    cv->eta = rmax(ETAMIN, cv->hmin / rabs(cv->h));
    cv->h *= cv->eta;
    cv->next_h = cv->h;
    cv->hscale = cv->h;
    cv->qwait = LONG_WAIT;
    cv->nscon = 0;
    const int retval = cv->f(cv->tn, cv->zn[0], cv->tempv, cv->user_data);
    cv->nfe++;
    if (retval != 0) return SHUD_CV_RHSFUNC_FAIL;
    N_VScale(cv->h, cv->tempv, cv->zn[1]);
    return TRY_AGAIN;
}

void complete_step(shud_cv *cv) {  // cvCompleteStep
    cv->nst++;
    cv->nscon++;
    cv->hu = cv->h;
    cv->qu = cv->q;
    for (int i = cv->q; i >= 2; i--) cv->tau[i] = cv->tau[i - 1];
    if (cv->q == 1 && cv->nst > 1) cv->tau[2] = cv->tau[1];
    cv->tau[1] = cv->h;
    if (cv->have_fused && cv->fused.complete_step && cv->abstol > 0.0 && cv->ewt_next &&
        cv->fused.complete_step(cv->fused.ctx, cv->q, cv->l, cv->acor, cv->zn, cv->reltol, cv->abstol, cv->ewt_next,
                                cv->yout_hint, &cv->nrm_next) == 0) {
        cv->ewt_next_valid = 1;
        cv->yout_done = cv->yout_hint != nullptr;
    } else {
        N_VScaleAddMulti(cv->q + 1, cv->l, cv->acor, cv->zn, cv->zn);
    }
    cv->qwait--;
    if (cv->qwait == 1 && cv->q != cv->qmax) {
        N_VScale(1.0, cv->acor, cv->zn[cv->qmax]);
        cv->saved_tq5 = cv->tq[5];
        cv->indx_acor = cv->qmax;
    }
}

void set_eta(shud_cv *cv) {  // cvSetEta
    if (cv->eta < THRESH) {
        cv->eta = 1.0;
        cv->hprime = cv->h;
    } else {
        cv->eta = rmin(cv->eta, cv->etamax);
        cv->eta /= rmax(1.0, rabs(cv->h) * cv->hmax_inv * cv->eta);
        cv->hprime = cv->h * cv->eta;
        if (cv->qprime < cv->q) cv->nscon = 0;
    }
}

void prepare_next_step(shud_cv *cv, double dsm) {  // cvPrepareNextStep
    if (cv->etamax == 1.0) {
        cv->qwait = cv->qwait > 2 ? cv->qwait : 2;
        cv->qprime = cv->q;
        cv->hprime = cv->h;
        cv->eta = 1.0;
        return;
    }
    const double etaq = 1.0 / (rpowerR(BIAS2 * dsm, 1.0 / cv->L) + ADDON);
    if (cv->qwait != 0) {
        cv->eta = etaq;
        cv->qprime = cv->q;
        set_eta(cv);
        return;
    }
    cv->qwait = 2;
    // cvComputeEtaqm1
    double etaqm1 = 0.0;
    if (cv->q > 1) {
        const double ddn = N_VWrmsNorm(cv->zn[cv->q], cv->ewt) * cv->tq[1];
        etaqm1 = 1.0 / (rpowerR(BIAS1 * ddn, 1.0 / cv->q) + ADDON);
    }
    // cvComputeEtaqp1
    double etaqp1 = 0.0;
    if (cv->q != cv->qmax && cv->saved_tq5 != 0.0) {
        const double cquot = (cv->tq[5] / cv->saved_tq5) * rpowerI(cv->h / cv->tau[2], cv->L);
        N_VLinearSum(-cquot, cv->zn[cv->qmax], 1.0, cv->acor, cv->tempv);
        const double dup = N_VWrmsNorm(cv->tempv, cv->ewt) * cv->tq[3];
        etaqp1 = 1.0 / (rpowerR(BIAS3 * dup, 1.0 / (cv->L + 1)) + ADDON);
    }
    // cvChooseEta
    const double etam = rmax(etaqm1, rmax(etaq, etaqp1));
    if (etam < THRESH) {
        cv->eta = 1.0;
        cv->qprime = cv->q;
    } else if (etam == etaq) {
        cv->eta = etaq;
        cv->qprime = cv->q;
    } else if (etam == etaqm1) {
        cv->eta = etaqm1;
        cv->qprime = cv->q - 1;
    } else {
        cv->eta = etaqp1;
        cv->qprime = cv->q + 1;
        N_VScale(1.0, cv->acor, cv->zn[cv->qmax]);  // Delta_n for the order increase
    }
    set_eta(cv);
}

int step(shud_cv *cv) {  // cvStep
    const double saved_t = cv->tn;
    int ncf = 0, nef = 0, nflag = FIRST_CALL;
    double dsm = 0.0;
    if (cv->nst > 0 && cv->hprime != cv->h) adjust_params(cv);
    for (;;) {
        predict(cv);
        set_coeffs(cv);
        nflag = nls(cv);
        const int kflag = handle_nflag(cv, &nflag, saved_t, &ncf);
        if (kflag == PREDICT_AGAIN) continue;
        if (kflag != DO_ERROR_TEST) return kflag;
        const int eflag = do_error_test(cv, &nflag, saved_t, &nef, &dsm);
        if (eflag == TRY_AGAIN) continue;
        if (eflag != 0) return eflag;
        break;
    }
    complete_step(cv);
    prepare_next_step(cv, dsm);
    cv->etamax = (cv->nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
    // acor * tq[2] is the estimated local error (CVodeGetEstLocalErrors); nothing of this integrator reads it and the
    // next step starts from acor = 0, so the device route leaves the pass out
    if (!(cv->have_fused && cv->fused.newton_step)) N_VScale(cv->tq[2], cv->acor, cv->acor);
    return 0;
}

// ---- initial step size (cvHin, cvUpperBoundH0, cvYddNorm) ----
double upper_bound_h0(shud_cv *cv, double tdist) {
    N_Vector temp1 = cv->tempv, temp2 = cv->acor;
    N_VAbs(cv->zn[0], temp2);
    ewt_set(cv, cv->zn[0], temp1);
    N_VInv(temp1, temp1);
    N_VLinearSum(HUB_FACTOR, temp2, 1.0, temp1, temp1);
    N_VAbs(cv->zn[1], temp2);
    N_VDiv(temp2, temp1, temp1);
    const double hub_inv = N_VMaxNorm(temp1);
    double hub = HUB_FACTOR * tdist;
    if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
    return hub;
}
int ydd_norm(shud_cv *cv, double hg, double *yddnrm) {
    N_VLinearSum(hg, cv->zn[1], 1.0, cv->zn[0], cv->y);
    const int retval = cv->f(cv->tn + hg, cv->y, cv->tempv, cv->user_data);
    cv->nfe++;
    if (retval < 0) return SHUD_CV_RHSFUNC_FAIL;
    if (retval > 0) return 1;
    N_VLinearSum(1.0 / hg, cv->tempv, -1.0 / hg, cv->zn[1], cv->tempv);
    *yddnrm = N_VWrmsNorm(cv->tempv, cv->ewt);
    return 0;
}
int hin(shud_cv *cv, double tout) {
    const int sign = (tout - cv->tn > 0.0) ? 1 : -1;
    const double tdist = rabs(tout - cv->tn);
    const double tround = cv->uround * rmax(rabs(cv->tn), rabs(tout));
    if (tdist < 2.0 * tround) return SHUD_CV_TOO_CLOSE;
    const double hlb = HLB_FACTOR * tround, hub = upper_bound_h0(cv, tdist);
    double hg = sqrt(hlb * hub);
    if (hub < hlb) { cv->h = sign == -1 ? -hg : hg; return 0; }
    bool hnewOK = false;
    double hs = hg, hnew = hg, yddnrm = 0.0;
    for (int count1 = 1; count1 <= MAX_HIN_ITERS; count1++) {
        bool hgOK = false;
        for (int count2 = 1; count2 <= MAX_HIN_ITERS; count2++) {
            const int retval = ydd_norm(cv, hg * sign, &yddnrm);
            if (retval < 0) return SHUD_CV_RHSFUNC_FAIL;
            if (retval == 0) { hgOK = true; break; }
            hg *= 0.2;
        }
        if (!hgOK) {
            if (count1 <= 2) return SHUD_CV_RHSFUNC_FAIL;
            hnew = hs;
            break;
        }
        hs = hg;
        if (hnewOK || count1 == MAX_HIN_ITERS) { hnew = hg; break; }
        hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
        const double hrat = hnew / hg;
        if (hrat > 0.5 && hrat < 2.0) hnewOK = true;
        if (count1 > 1 && hrat > 2.0) { hnew = hg; hnewOK = true; }
        hg = hnew;
    }
    double h0 = H_BIAS * hnew;
    if (h0 < hlb) h0 = hlb;
    if (h0 > hub) h0 = hub;
    cv->h = sign == -1 ? -h0 : h0;
    return 0;
}

void free_vec(N_Vector *v) { if (*v) { N_VDestroy(*v); *v = nullptr; } }

}  // namespace

extern "C" {

int shud_cv_create(shud_cv_rhs_fn f, void *user_data, realtype t0, N_Vector y0, shud_cv **out) {
    if (!f || !y0 || !out) return SHUD_CV_ILL_INPUT;
    shud_cv *cv = (shud_cv *)calloc(1, sizeof(shud_cv));
    if (!cv) return SHUD_CV_MEM_FAIL;
    cv->f = f; cv->user_data = user_data;
    cv->uround = DBL_EPSILON;
    cv->qmax = Q_MAX; cv->maxl = SPGMR_MAXL_DEFAULT;
    cv->mxstep = 500;  // MXSTEP_DEFAULT
    cv->nlscoef = NLSCOEF;
    cv->hmin = 0.0; cv->hmax_inv = 0.0; cv->hin = 0.0;
    bool ok = true;
    for (int j = 0; j <= L_MAX; j++) ok = ok && (cv->zn[j] = N_VClone(y0));
    ok = ok && (cv->ewt_next = N_VClone(y0));
    ok = ok && (cv->ewt = N_VClone(y0)) && (cv->y = N_VClone(y0)) && (cv->acor = N_VClone(y0)) &&
         (cv->tempv = N_VClone(y0)) && (cv->ftemp = N_VClone(y0)) && (cv->vtemp1 = N_VClone(y0));
    for (int k = 0; k <= SPGMR_MAXL_DEFAULT; k++) ok = ok && (cv->V[k] = N_VClone(y0));
    ok = ok && (cv->xcor = N_VClone(y0)) && (cv->ls_x = N_VClone(y0)) && (cv->ls_vtemp = N_VClone(y0));
    if (!ok) { shud_cv_free(cv); return SHUD_CV_MEM_FAIL; }
    cv->nrmfac = sqrt((double)N_VGetLength(y0));  // CVLS: norm conversion factor sqrt(N)
    *out = cv;
    return shud_cv_reinit(cv, t0, y0);
}

void shud_cv_free(shud_cv *cv) {
    if (!cv) return;
    for (int j = 0; j <= L_MAX; j++) free_vec(&cv->zn[j]);
    free_vec(&cv->ewt_next);
    free_vec(&cv->ewt); free_vec(&cv->y); free_vec(&cv->acor); free_vec(&cv->tempv); free_vec(&cv->ftemp); free_vec(&cv->vtemp1);
    for (int k = 0; k <= SPGMR_MAXL_DEFAULT; k++) free_vec(&cv->V[k]);
    free_vec(&cv->xcor); free_vec(&cv->ls_x); free_vec(&cv->ls_vtemp);
    free(cv);
}

int shud_cv_reinit(shud_cv *cv, realtype t0, N_Vector y0) {
    if (!cv || !y0) return SHUD_CV_ILL_INPUT;
    cv->tn = t0;
    cv->q = 1; cv->L = 2; cv->qwait = cv->L; cv->etamax = ETAMX1;
    cv->qu = 0; cv->hu = 0.0; cv->tretlast = t0;
    N_VScale(1.0, y0, cv->zn[0]);
    cv->nst = cv->nfe = cv->ncfn = cv->netf = cv->nni = cv->nscon = 0;
    cv->nli = cv->ncfl = cv->nfeLS = cv->nps = 0;
    cv->h0u = 0.0; cv->next_h = 0.0; cv->next_q = 0; cv->h = 0.0;
    cv->saved_tq5 = 0.0; cv->indx_acor = 0; cv->nls_primed = 0; cv->ewt_next_valid = 0; cv->yout_done = 0;
    memset(cv->tau, 0, sizeof(cv->tau)); memset(cv->tq, 0, sizeof(cv->tq)); memset(cv->l, 0, sizeof(cv->l));
    return SHUD_CV_SUCCESS;
}

int shud_cv_sstolerances(shud_cv *cv, realtype reltol, realtype abstol) {
    if (!cv || reltol < 0.0 || abstol < 0.0) return SHUD_CV_ILL_INPUT;
    cv->reltol = reltol; cv->abstol = abstol; cv->tol_set = 1;
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_max_ord(shud_cv *cv, int maxord) {
    if (!cv || maxord < 1 || maxord > Q_MAX) return SHUD_CV_ILL_INPUT;
    cv->qmax = maxord;
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_min_step(shud_cv *cv, realtype hmin) {
    if (!cv || hmin < 0.0) return SHUD_CV_ILL_INPUT;
    if (hmin * cv->hmax_inv > 1.0) return SHUD_CV_ILL_INPUT;
    cv->hmin = hmin;
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_max_step(shud_cv *cv, realtype hmax) {
    if (!cv || hmax < 0.0) return SHUD_CV_ILL_INPUT;
    if (hmax == 0.0) { cv->hmax_inv = 0.0; return SHUD_CV_SUCCESS; }
    const double hmax_inv = 1.0 / hmax;
    if (hmax_inv * cv->hmin > 1.0) return SHUD_CV_ILL_INPUT;
    cv->hmax_inv = hmax_inv;
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_init_step(shud_cv *cv, realtype hin_) { if (!cv) return SHUD_CV_ILL_INPUT; cv->hin = hin_; return SHUD_CV_SUCCESS; }
int shud_cv_set_max_num_steps(shud_cv *cv, long mxsteps) {
    if (!cv) return SHUD_CV_ILL_INPUT;
    cv->mxstep = mxsteps == 0 ? 500 : mxsteps;  // 0: default; negative: no limit
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_stop_time(shud_cv *cv, realtype tstop) {
    if (!cv) return SHUD_CV_ILL_INPUT;
    if (cv->nst > 0 && (tstop - cv->tn) * cv->h < 0.0) return SHUD_CV_ILL_INPUT;
    cv->tstop = tstop; cv->tstopset = 1;
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_maxl(shud_cv *cv, int maxl) {
    if (!cv || maxl > SPGMR_MAXL_DEFAULT) return SHUD_CV_ILL_INPUT;
    cv->maxl = maxl <= 0 ? SPGMR_MAXL_DEFAULT : maxl;
    return SHUD_CV_SUCCESS;
}
int shud_cv_set_fused(shud_cv *cv, const shud_cv_fused *fused) {
    if (!cv) return SHUD_CV_ILL_INPUT;
    if (fused) { cv->fused = *fused; cv->have_fused = 1; }
    else cv->have_fused = 0;
    return SHUD_CV_SUCCESS;
}

int shud_cv_get_dky(shud_cv *cv, realtype t, int k, N_Vector dky) {  // CVodeGetDky
    if (!cv || !dky || k < 0 || k > cv->q) return SHUD_CV_ILL_INPUT;
    double tfuzz = FUZZ_FACTOR * cv->uround * (rabs(cv->tn) + rabs(cv->hu));
    if (cv->hu < 0.0) tfuzz = -tfuzz;
    const double tp = cv->tn - cv->hu - tfuzz, tn1 = cv->tn + tfuzz;
    if ((t - tp) * (t - tn1) > 0.0) return SHUD_CV_BAD_T;
    const double s = (t - cv->tn) / cv->h;
    double cvals[L_MAX + 1];
    N_Vector X[L_MAX + 1];
    int nvec = 0;
    for (int j = cv->q; j >= k; j--) {
        double c = 1.0;
        for (int i = j; i >= j - k + 1; i--) c *= i;
        for (int i = 0; i < j - k; i++) c *= s;
        cvals[nvec] = c; X[nvec] = cv->zn[j]; nvec++;
    }
    if (N_VLinearCombination(nvec, cvals, X, dky) != 0) return SHUD_CV_MEM_FAIL;
    if (k == 0) return SHUD_CV_SUCCESS;
    N_VScale(rpowerI(cv->h, -k), dky, dky);
    return SHUD_CV_SUCCESS;
}

int shud_cv_solve(shud_cv *cv, realtype tout, N_Vector yout, realtype *tret, int itask) {  // CVode
    if (!cv || !yout || !tret || !cv->tol_set) return SHUD_CV_ILL_INPUT;
    if (itask != SHUD_CV_NORMAL && itask != SHUD_CV_ONE_STEP) return SHUD_CV_ILL_INPUT;
    double troundoff;
    if (cv->nst == 0) {
        // first call: cvInitialSetup, f(t0, y0), initial step
        cv->tretlast = *tret = cv->tn;
        if (ewt_set(cv, cv->zn[0], cv->ewt) != 0) return SHUD_CV_ILL_INPUT;
        const int retval = cv->f(cv->tn, cv->zn[0], cv->zn[1], cv->user_data);
        cv->nfe++;
        if (retval != 0) return SHUD_CV_RHSFUNC_FAIL;
        if (cv->tstopset) {
            if ((cv->tstop - cv->tn) * (tout - cv->tn) <= 0.0) return SHUD_CV_ILL_INPUT;
        }
        cv->h = cv->hin;
        if (cv->h != 0.0 && (tout - cv->tn) * cv->h < 0.0) return SHUD_CV_ILL_INPUT;
        if (cv->h == 0.0) {
            double tout_hin = tout;
            if (cv->tstopset && (tout - cv->tn) * (tout - cv->tstop) > 0.0) tout_hin = cv->tstop;
            const int hflag = hin(cv, tout_hin);
            if (hflag != 0) return hflag;
        }
        const double rh = rabs(cv->h) * cv->hmax_inv;
        if (rh > 1.0) cv->h /= rh;
        if (rabs(cv->h) < cv->hmin) cv->h *= cv->hmin / rabs(cv->h);
        if (cv->tstopset) {
            if ((cv->tn + cv->h - cv->tstop) * cv->h > 0.0) cv->h = (cv->tstop - cv->tn) * (1.0 - 4.0 * cv->uround);
        }
        cv->hscale = cv->h; cv->h0u = cv->h; cv->hprime = cv->h;
        N_VScale(cv->h, cv->zn[1], cv->zn[1]);
    } else {
        // later calls: stop tests before stepping
        troundoff = FUZZ_FACTOR * cv->uround * (rabs(cv->tn) + rabs(cv->h));
        if (itask == SHUD_CV_NORMAL && (cv->tn - tout) * cv->h >= 0.0) {
            cv->tretlast = *tret = tout;
            return shud_cv_get_dky(cv, tout, 0, yout) != 0 ? SHUD_CV_ILL_INPUT : SHUD_CV_SUCCESS;
        }
        if (itask == SHUD_CV_ONE_STEP && rabs(cv->tn - cv->tretlast) > troundoff) {
            cv->tretlast = *tret = cv->tn;
            N_VScale(1.0, cv->zn[0], yout);
            return SHUD_CV_SUCCESS;
        }
        if (cv->tstopset) {
            if (rabs(cv->tn - cv->tstop) <= troundoff) {
                if (shud_cv_get_dky(cv, cv->tstop, 0, yout) != 0) return SHUD_CV_ILL_INPUT;
                cv->tretlast = *tret = cv->tstop;
                cv->tstopset = 0;
                return SHUD_CV_TSTOP_RETURN;
            }
            if ((cv->tn + cv->hprime - cv->tstop) * cv->h > 0.0) {
                cv->hprime = (cv->tstop - cv->tn) * (1.0 - 4.0 * cv->uround);
                cv->eta = cv->hprime / cv->h;
            }
        }
    }
    long nstloc = 0;
    int istate = SHUD_CV_SUCCESS;
    cv->yout_hint = itask == SHUD_CV_ONE_STEP ? yout : nullptr;
    cv->yout_done = 0;
    for (;;) {
        cv->next_h = cv->h;
        cv->next_q = cv->q;
        double nrm = -1.0;
        if (cv->nst > 0) {
            if (cv->ewt_next_valid) {
                // formed by the fused cvCompleteStep of the step before, from the zn[0] that is still there
                N_Vector tmp = cv->ewt; cv->ewt = cv->ewt_next; cv->ewt_next = tmp;
                nrm = cv->nrm_next;
                cv->ewt_next_valid = 0;
            } else if (cv->have_fused && cv->fused.ewt_set_norm && cv->abstol > 0.0) {
                // weights and the norm of the tolsf test below in one pass over zn[0]
                if (cv->fused.ewt_set_norm(cv->fused.ctx, cv->reltol, cv->abstol, cv->zn[0], cv->ewt, &nrm) != 0) nrm = -1.0;
            }
            if (nrm < 0.0 && ewt_set(cv, cv->zn[0], cv->ewt) != 0) {
                istate = SHUD_CV_ILL_INPUT;
                cv->tretlast = *tret = cv->tn;
                N_VScale(1.0, cv->zn[0], yout);
                break;
            }
        }
        if (cv->mxstep > 0 && nstloc >= cv->mxstep) {
            istate = SHUD_CV_TOO_MUCH_WORK;
            cv->tretlast = *tret = cv->tn;
            N_VScale(1.0, cv->zn[0], yout);
            break;
        }
        if (nrm < 0.0) nrm = N_VWrmsNorm(cv->zn[0], cv->ewt);
        if (cv->uround * nrm > 1.0) {
            istate = SHUD_CV_TOO_MUCH_ACC;
            cv->tretlast = *tret = cv->tn;
            N_VScale(1.0, cv->zn[0], yout);
            break;
        }
        const int kflag = step(cv);
        if (kflag != 0) {
            istate = kflag;
            cv->tretlast = *tret = cv->tn;
            N_VScale(1.0, cv->zn[0], yout);
            break;
        }
        nstloc++;
        if (itask == SHUD_CV_NORMAL && (cv->tn - tout) * cv->h >= 0.0) {
            istate = SHUD_CV_SUCCESS;
            cv->tretlast = *tret = tout;
            shud_cv_get_dky(cv, tout, 0, yout);
            cv->next_q = cv->qprime; cv->next_h = cv->hprime;
            break;
        }
        if (cv->tstopset) {
            troundoff = FUZZ_FACTOR * cv->uround * (rabs(cv->tn) + rabs(cv->h));
            if (rabs(cv->tn - cv->tstop) <= troundoff) {
                shud_cv_get_dky(cv, cv->tstop, 0, yout);
                cv->tretlast = *tret = cv->tstop;
                cv->tstopset = 0;
                istate = SHUD_CV_TSTOP_RETURN;
                break;
            }
            if ((cv->tn + cv->hprime - cv->tstop) * cv->h > 0.0) {
                cv->hprime = (cv->tstop - cv->tn) * (1.0 - 4.0 * cv->uround);
                cv->eta = cv->hprime / cv->h;
            }
        }
        if (itask == SHUD_CV_ONE_STEP) {
            istate = SHUD_CV_SUCCESS;
            cv->tretlast = *tret = cv->tn;
            if (!cv->yout_done) N_VScale(1.0, cv->zn[0], yout);
            cv->next_q = cv->qprime; cv->next_h = cv->hprime;
            break;
        }
    }
    return istate;
}

// One linear solve (I - gamma J(t, y)) x = b on its own: SUNLinSolSolve_SPGMR as CVLS drives it (scaling by `ewt` on both
// sides, zero initial guess, difference-quotient J v around (y, fy)), through the fused hook when one is set.
// Returns 0 converged, 1 residual reduced, 2 not reduced, 3 right-hand side below delta (x = 0), < 0 error.
int shud_cv_linsolve(shud_cv *cv, realtype t, realtype gamma, N_Vector y, N_Vector fy, N_Vector ewt, N_Vector b,
                     realtype delta, N_Vector x, int *nli) {
    if (!cv || !y || !fy || !ewt || !b || !x) return SHUD_CV_ILL_INPUT;
    cv->tn = t; cv->gamma = gamma;
    N_VScale(1.0, ewt, cv->ewt);
    int it = 0, nfe = 0, r;
    if (cv->have_fused && cv->fused.lsolve) {
        r = cv->fused.lsolve(cv->fused.ctx, t, gamma, y, fy, cv->ewt, b, delta, x, &it, &nfe);
    } else {
        const int rc = spgmr_solve(cv, x, b, delta, y, fy, &it);
        r = rc == LS_SUCCESS ? (it == 0 ? 3 : 0) : rc == LS_RES_REDUCED ? 1 : rc == LS_CONV_FAIL ? 2 : -1;
    }
    if (nli) *nli = it;
    return r;
}

int shud_cv_get_stats(const shud_cv *cv, shud_cv_stats *st) {
    if (!cv || !st) return SHUD_CV_ILL_INPUT;
    st->nst = cv->nst; st->nfe = cv->nfe; st->nfeLS = cv->nfeLS; st->nni = cv->nni; st->nli = cv->nli;
    st->ncfn = cv->ncfn; st->netf = cv->netf; st->ncfl = cv->ncfl;
    st->qlast = cv->qu; st->qcur = cv->next_q; st->hinused = cv->h0u; st->hlast = cv->hu; st->hcur = cv->next_h;
    st->tcur = cv->tn;
    return SHUD_CV_SUCCESS;
}

}  // extern "C"
