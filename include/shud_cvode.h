/* shud_cvode.h - the integrator side of the boundary: a restatement of the CVODE configuration the reference runs,
 * written ONLY against the generic N_Vector operations table (v->ops->...), so that the same object code drives
 *   - the device N_Vector + shud_b200_f (GPU arm), and
 *   - a host serial N_Vector + the CPU oracle RHS (checker arm, tests only),
 * and so that the N_Vector table of shud_sundials.h is exercised the way CVODE exercises it.
 *
 * SUNDIALS CVODE is third-party and absent from the reference tree and from this image (configure:17-21 pins
 * cvode-6.0.0).  This file restates the published CVODE 6 algorithm for exactly the configuration of
 * src/Equations/cvode_config.cpp:162-193 and src/Model/shud.cpp:89-133:
 *   CVodeCreate(CV_BDF): variable-order (1..5) variable-step BDF in fixed-leading-coefficient Nordsieck form;
 *   CVodeSStolerances(reltol, abstol); CVodeSetMinStep / SetMaxStep / SetInitStep / SetMaxNumSteps / SetStopTime;
 *   Newton iteration (<= 3 corrector iterations, nlscoef 0.1, convergence-rate estimate crdown 0.3, rdiv 2);
 *   SUNLinSol_SPGMR(PREC_NONE, maxl 5, modified Gram-Schmidt with the re-orthogonalisation test, no restarts) driven
 *   matrix-free by CVLS: difference-quotient J v with sigma = 1 / ||v||_WRMS, eplifac 0.05, norm factor sqrt(N),
 *   ewt scaling on both sides, a reduced residual accepted on the first Newton iteration only;
 *   step / order selection (eta_q, eta_q-1, eta_q+1 with biases 6 / 6 / 10, threshold 1.5, eta_max 10000 / 10 / 10),
 *   error-test and convergence-failure handling (eta 0.25 on convergence failure, order reduction after 3 error-test
 *   failures, at most 7 / 10 failures per step), dense output (CVodeGetDky).
 * Not restated (unused by the reference): Adams, root finding, constraints, projection, sensitivity, stability-limit
 * detection, preconditioning, matrix-based linear solvers, nonlinear solvers other than Newton.
 * "parity unpinned": no reference test pins step sequences; both arms of every comparison run THIS code.
 */
#ifndef SHUD_CVODE_H
#define SHUD_CVODE_H
#include "shud_sundials.h"
#ifdef __cplusplus
extern "C" {
#endif

/* return flags (CVODE's values) */
#define SHUD_CV_SUCCESS 0
#define SHUD_CV_TSTOP_RETURN 1
#define SHUD_CV_TOO_MUCH_WORK (-1)
#define SHUD_CV_TOO_MUCH_ACC (-2)
#define SHUD_CV_ERR_FAILURE (-3)
#define SHUD_CV_CONV_FAILURE (-4)
#define SHUD_CV_LSOLVE_FAIL (-7)
#define SHUD_CV_RHSFUNC_FAIL (-8)
#define SHUD_CV_MEM_FAIL (-20)
#define SHUD_CV_ILL_INPUT (-22)
#define SHUD_CV_BAD_T (-25)
#define SHUD_CV_TOO_CLOSE (-27)
#define SHUD_CV_NORMAL 1
#define SHUD_CV_ONE_STEP 2

typedef int (*shud_cv_rhs_fn)(realtype t, N_Vector y, N_Vector ydot, void *user_data); /* CVRhsFn */
typedef struct shud_cv shud_cv;

/* Optional fused vector operations of the integrator.  Any member may be NULL; the generic
 * ops-table sequence is used then.  Each must compute exactly what the comment says (same operation order). */
typedef struct shud_cv_fused {
    void *ctx;
    /* ewt = 1 ./ (rtol |y| + atol)                                     (cvEwtSetSS: Abs, Scale, AddConst, Inv) */
    int (*ewt_set)(void *ctx, realtype rtol, realtype atol, N_Vector y, N_Vector ewt);
    /* res = rl1 zn1 + ycor - gamma f                                   (cvNlsResidual's two LinearSums) */
    int (*nls_residual)(void *ctx, realtype rl1, N_Vector zn1, N_Vector ycor, realtype gamma, N_Vector f, N_Vector res);
    /* the whole linear solve (I - gamma J) x = b, J v by difference quotients of `f` around (t, y, fy), scaled by ewt,
     * zero initial guess, tolerance delta on the scaled residual 2-norm; returns 0 converged, 1 residual reduced,
     * 2 not reduced, 3 the scaled right-hand side was already below delta (nothing done: cvLsSolve's norm test),
     * < 0 error; *nli Krylov iterations, *nfe RHS evaluations spent */
    int (*lsolve)(void *ctx, realtype t, realtype gamma, N_Vector y, N_Vector fy, N_Vector ewt, N_Vector b, realtype delta,
                  N_Vector x, int *nli, int *nfe);
    /* cvPredict (sgn = +1) / cvRestore (sgn = -1): zn[j-1] += sgn zn[j] for k = 1..q, j = q..k, in place; with
     * acor != NULL also the start of the Newton iteration: acor = 0, y = zn[0] + acor */
    int (*predict)(void *ctx, int q, realtype sgn, N_Vector *zn, N_Vector y, N_Vector acor);
    /* one Newton iteration around y (fy = f(t, y)): solve (I - gamma J) x = -(rl1 zn1 + acor - gamma fy) as lsolve
     * does, then acor += x, y = zn0 + acor, *del = ||x||_WRMS(ewt); returns lsolve's codes - 3: nothing was written,
     * the caller runs the iteration through nls_residual / lsolve */
    int (*newton_step)(void *ctx, realtype t, realtype gamma, realtype rl1, N_Vector zn0, N_Vector zn1, N_Vector acor,
                       N_Vector y, N_Vector fy, N_Vector ewt, realtype delta, realtype *del, int *nli, int *nfe);
    /* ewt_set and *nrm = ||y||_WRMS(ewt) in one pass (the weights and the tolsf test at the top of CVode's loop) */
    int (*ewt_set_norm)(void *ctx, realtype rtol, realtype atol, N_Vector y, N_Vector ewt, realtype *nrm);
    /* cvCompleteStep's zn[j] += l[j] acor (j = 0..q) together with what the top of CVode()'s loop does next with the
     * new zn[0]: ewt_next = 1 ./ (rtol |zn0| + atol), *nrm = ||zn0||_WRMS(ewt_next), and yout = zn0 when yout != NULL
     * (CV_ONE_STEP's copy).  ewt_next is a second weight vector: the step's own weights stay valid for the order
     * selection that follows (cvPrepareNextStep). */
    int (*complete_step)(void *ctx, int q, realtype *l, N_Vector acor, N_Vector *zn, realtype rtol, realtype atol,
                         N_Vector ewt_next, N_Vector yout, realtype *nrm);
} shud_cv_fused;

/* CVodeCreate(CV_BDF) + CVodeInit(f, t0, y0) + CVodeSetUserData: work vectors are cloned from y0 */
int shud_cv_create(shud_cv_rhs_fn f, void *user_data, realtype t0, N_Vector y0, shud_cv **out);
void shud_cv_free(shud_cv *cv);                                       /* CVodeFree */
int shud_cv_reinit(shud_cv *cv, realtype t0, N_Vector y0);            /* CVodeReInit */
int shud_cv_sstolerances(shud_cv *cv, realtype reltol, realtype abstol);
int shud_cv_set_max_ord(shud_cv *cv, int maxord);                      /* 1..5, default 5 */
int shud_cv_set_min_step(shud_cv *cv, realtype hmin);
int shud_cv_set_max_step(shud_cv *cv, realtype hmax);
int shud_cv_set_init_step(shud_cv *cv, realtype hin);
int shud_cv_set_max_num_steps(shud_cv *cv, long mxsteps);
int shud_cv_set_stop_time(shud_cv *cv, realtype tstop);
int shud_cv_set_maxl(shud_cv *cv, int maxl);                           /* SUNLinSol_SPGMR(y, PREC_NONE, maxl): 0 -> 5 */
int shud_cv_set_fused(shud_cv *cv, const shud_cv_fused *fused);
/* CVode(mem, tout, yout, &tret, itask) */
int shud_cv_solve(shud_cv *cv, realtype tout, N_Vector yout, realtype *tret, int itask);
int shud_cv_get_dky(shud_cv *cv, realtype t, int k, N_Vector dky);     /* CVodeGetDky */
/* statistics: nst, nfe (+ nfeLS), nni, nli, ncfn, netf, ncfl, qlast, qcur, hlast, hcur, tcur (CVodeGet*) */
typedef struct shud_cv_stats {
    long nst, nfe, nfeLS, nni, nli, ncfn, netf, ncfl;
    int qlast, qcur;
    realtype hinused, hlast, hcur, tcur;
} shud_cv_stats;
int shud_cv_get_stats(const shud_cv *cv, shud_cv_stats *st);

/* One linear solve (I - gamma J(t, y)) x = b on its own - SUNLinSolSolve_SPGMR as CVLS drives it: scaling by `ewt` on
 * both sides, zero initial guess, difference-quotient J v around (y, fy = f(t, y)), tolerance `delta` on the 2-norm of
 * the scaled residual; runs the fused hook when one is set.  Returns 0 converged, 1 residual reduced, 2 not reduced,
 * 3 right-hand side already below delta (x = 0), < 0 error; *nli = Krylov iterations. */
int shud_cv_linsolve(shud_cv *cv, realtype t, realtype gamma, N_Vector y, N_Vector fy, N_Vector ewt, N_Vector b,
                     realtype delta, N_Vector x, int *nli);

/* The device implementation of shud_cv_fused for one GPU (shud_nv_ewt, one-pass Newton residual, shud_spgmr_solve with
 * the difference-quotient work folded around shud_b200_rhs_dev): fills *out; destroy releases its SPGMR workspace.
 * The vectors handed to the hooks must be SHUD B200 vectors on `ws`. */
struct shud_ctx;
int shud_b200_cv_fused_create(struct shud_ctx *gpu, shud_nvws *ws, int maxl, shud_cv_fused *out);
void shud_b200_cv_fused_destroy(shud_cv_fused *f);

#ifdef __cplusplus
}
#endif
#endif
