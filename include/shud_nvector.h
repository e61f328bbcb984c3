/* shud_nvector.h - C ABI of the device N_Vector arithmetic CVODE(BDF/Newton) + CVLS + SPGMR run on
 * the SHUD state vectors (SURVEY.md section 8(a) row a22, 8(b) "N_Vector ops table").
 *
 * The reference takes these from SUNDIALS' nvector_serial / nvector_openmp (linked at reference
 * Makefile:87-88, vectors created at src/Model/shud.cpp:59-64, cloned by CVODE and SPGMR through
 * src/Equations/cvode_config.cpp:169,176).  SUNDIALS is not vendored in the reference; the semantics
 * below are those of the SUNDIALS 6 N_Vector documentation (the version the reference pins:
 * configure:17-21, README.md:18).  Each function names the ops-table member it replaces.
 *
 * All vector arguments are DEVICE pointers to `n` contiguous doubles.  Streaming operations are
 * asynchronous on the workspace's stream; reductions return their value through a host pointer and
 * synchronise the stream (CVODE consumes every norm / dot product on the host).  Reductions use a
 * fixed grid and a fixed-shape warp-shuffle tree: deterministic run to run.  There is no CPU fallback.
 */
#ifndef SHUD_NVECTOR_H
#define SHUD_NVECTOR_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SHUD_NV_MAXVEC 8 /* vectors per fused call (CVODE: q+1 <= 6 Nordsieck rows; SPGMR: maxl+1 = 6) */

typedef struct shud_nvws shud_nvws; /* reduction workspace bound to one device + stream */

/* `stream` is a cudaStream_t (e.g. shud_b200_stream()); NULL = the legacy default stream */
int shud_nv_ws_create(int device, void *stream, shud_nvws **out);
void shud_nv_ws_destroy(shud_nvws *ws);
void *shud_nv_ws_stream(const shud_nvws *ws); /* the cudaStream_t the workspace was created on */
/* Distributed vectors (one partition per GPU): with an allreduce installed every reduction below returns the GLOBAL
 * value - the per-rank partials stay in device memory, `fn(ctx, dev_vals, n, op, stream)` reduces them over the ranks
 * in place on the workspace's stream (op 0 sum, 1 max, 2 min; shud_b200_allreduce_dev = ncclAllReduce over the context's
 * communicator), and only the result crosses to the host: one synchronisation per reduction.  Every rank must issue the
 * same reductions.  shud_nv_ws_local(ws, 1) ... shud_nv_ws_local(ws, 0) brackets calls that must stay local. */
typedef int (*shud_nv_allreduce_dev_fn)(void *ctx, double *dev_vals, int n, int op, void *stream);
int shud_nv_ws_set_allreduce(shud_nvws *ws, shud_nv_allreduce_dev_fn fn, void *ctx);
/* Allreduce INSIDE the reduction kernels: boxes[r] = device pointer to rank r's mailbox (SHUD_NV_ARBOX_BYTES of zeroed
 * device memory, the other ranks' mapped through CUDA IPC - shud_b200_p2p_mailboxes hands them out).  The last block of
 * every reduction stores the rank's partials into every mailbox over NVLink, waits for the others' and combines them in
 * rank order: one kernel per global reduction, the same bits on every rank.  Takes precedence over the hook above;
 * one workspace per set of mailboxes; every rank must issue the same reductions in the same order.  nranks <= 1: off. */
#define SHUD_NV_MAXRANKS 16
#define SHUD_NV_ARBOX_BYTES 4096
int shud_nv_ws_set_peer_allreduce(shud_nvws *ws, int nranks, int rank, void *const *boxes);
void shud_nv_ws_local(shud_nvws *ws, int on);
int shud_nv_ws_device(const shud_nvws *ws);

/* ---- streaming operations (return 0 or a negative SHUD_ERR_* code) ---- */
int shud_nv_linearsum(shud_nvws *ws, int64_t n, double a, const double *x, double b, const double *y, double *z); /* nvlinearsum  z = a x + b y */
int shud_nv_const(shud_nvws *ws, int64_t n, double c, double *z);                                                /* nvconst      z = c */
int shud_nv_prod(shud_nvws *ws, int64_t n, const double *x, const double *y, double *z);                        /* nvprod       z = x .* y */
int shud_nv_div(shud_nvws *ws, int64_t n, const double *x, const double *y, double *z);                         /* nvdiv        z = x ./ y */
int shud_nv_scale(shud_nvws *ws, int64_t n, double c, const double *x, double *z);                              /* nvscale      z = c x */
int shud_nv_abs(shud_nvws *ws, int64_t n, const double *x, double *z);                                          /* nvabs        z = |x| */
int shud_nv_inv(shud_nvws *ws, int64_t n, const double *x, double *z);                                          /* nvinv        z = 1 ./ x */
int shud_nv_addconst(shud_nvws *ws, int64_t n, const double *x, double b, double *z);                           /* nvaddconst   z = x + b */
int shud_nv_compare(shud_nvws *ws, int64_t n, double c, const double *x, double *z);                            /* nvcompare    z = |x| >= c ? 1 : 0 */

/* ---- reductions (value in *out on the host; the call synchronises the stream) ---- */
int shud_nv_dotprod(shud_nvws *ws, int64_t n, const double *x, const double *y, double *out);    /* nvdotprod (+ nvdotprodlocal) */
int shud_nv_maxnorm(shud_nvws *ws, int64_t n, const double *x, double *out);                     /* nvmaxnorm (+ nvmaxnormlocal) */
int shud_nv_min(shud_nvws *ws, int64_t n, const double *x, double *out);                         /* nvmin     (+ nvminlocal)     */
int shud_nv_l1norm(shud_nvws *ws, int64_t n, const double *x, double *out);                      /* nvl1norm  */
int shud_nv_wsqrsum(shud_nvws *ws, int64_t n, const double *x, const double *w, double *out);    /* nvwsqrsumlocal  sum (x w)^2 */
int shud_nv_wsqrsum_mask(shud_nvws *ws, int64_t n, const double *x, const double *w, const double *id, double *out); /* nvwsqrsummasklocal */
/* nvwrmsnorm = sqrt(wsqrsum / n_global); nvwl2norm = sqrt(wsqrsum).  n_global is the length of the whole
 * (possibly distributed) vector: a single-GPU caller passes n. */
int shud_nv_wrmsnorm(shud_nvws *ws, int64_t n, const double *x, const double *w, int64_t n_global, double *out);
int shud_nv_wrmsnorm_mask(shud_nvws *ws, int64_t n, const double *x, const double *w, const double *id, int64_t n_global, double *out);
int shud_nv_wl2norm(shud_nvws *ws, int64_t n, const double *x, const double *w, double *out);
/* nvinvtest: z = 1./x where x != 0; *all_nonzero = 1 if no zero was met.  nvconstrmask: SUNDIALS' constraint
 * test (c = +-1: x c >= 0 ... , +-2: x c > 0), m = 1 where violated; *all_ok = 1 if none.
 * nvminquotient: min over denom != 0 of num/denom (DBL_MAX if none). */
int shud_nv_invtest(shud_nvws *ws, int64_t n, const double *x, double *z, int *all_nonzero);
int shud_nv_constrmask(shud_nvws *ws, int64_t n, const double *c, const double *x, double *m, int *all_ok);
int shud_nv_minquotient(shud_nvws *ws, int64_t n, const double *num, const double *denom, double *out);

/* ---- fused operations (pointer arrays are HOST arrays of device pointers, nvec <= SHUD_NV_MAXVEC) ---- */
int shud_nv_linearcombination(shud_nvws *ws, int64_t n, int nvec, const double *c, const double *const *X, double *z);         /* nvlinearcombination  z = sum c_k X_k */
int shud_nv_scaleaddmulti(shud_nvws *ws, int64_t n, int nvec, const double *a, const double *x, const double *const *Y, double *const *Z); /* nvscaleaddmulti  Z_k = a_k x + Y_k */
int shud_nv_dotprodmulti(shud_nvws *ws, int64_t n, int nvec, const double *x, const double *const *Y, double *out);            /* nvdotprodmulti  out_k = x . Y_k */
int shud_nv_linearsumvectorarray(shud_nvws *ws, int64_t n, int nvec, double a, const double *const *X, double b, const double *const *Y, double *const *Z);
int shud_nv_scalevectorarray(shud_nvws *ws, int64_t n, int nvec, const double *c, const double *const *X, double *const *Z);
int shud_nv_constvectorarray(shud_nvws *ws, int64_t n, int nvec, double c, double *const *Z);
int shud_nv_wrmsnormvectorarray(shud_nvws *ws, int64_t n, int nvec, const double *const *X, const double *const *W, int64_t n_global, double *out);

/* ---- integrator-level fusions: the vector work CVLS + SPGMR wrap around each
 * difference-quotient Jacobian-vector product, one launch each instead of two / three ops-table calls ----
 * shud_nv_dq_perturb:  ytemp = y + sigma * (vs ./ ewt)          (unscale the Krylov vector, perturb the state)
 * shud_nv_dq_combine:  out = ewt .* ( (vs ./ ewt) - gamma * (fpert - fy) / sigma )
 *                      = scaled (I - gamma J) applied to the unscaled Krylov vector, J v by difference quotient */
int shud_nv_dq_perturb(shud_nvws *ws, int64_t n, double sigma, const double *vs, const double *ewt, const double *y, double *ytemp);
int shud_nv_dq_combine(shud_nvws *ws, int64_t n, double sigma, double gamma, const double *vs, const double *ewt,
                       const double *fpert, const double *fy, double *out);

/* shud_nv_ewt:           ewt = 1 ./ (rtol |y| + atol)                 (CVODE's cvEwtSetSS: Abs, Scale, AddConst, Inv)
 * shud_nv_newton_resid:  r = gamma f + psi - y                        (right-hand side of the Newton system)
 * shud_nv_newton_update: y += x; acor += x; *del = ||x||_WRMS(ewt)    (Newton correction + its convergence norm) */
int shud_nv_ewt(shud_nvws *ws, int64_t n, double rtol, double atol, const double *y, double *ewt);
/* ewt = 1 ./ (rtol |y| + atol) and *nrm = ||y||_WRMS(ewt) in one pass (cvEwtSet + the tolsf test of CVode's loop) */
int shud_nv_ewt_wrms(shud_nvws *ws, int64_t n, double rtol, double atol, const double *y, double *ewt, int64_t n_global,
                     double *nrm);
int shud_nv_newton_resid(shud_nvws *ws, int64_t n, double gamma, const double *f, const double *psi, const double *y, double *r);
int shud_nv_newton_update(shud_nvws *ws, int64_t n, const double *x, const double *ewt, int64_t n_global, double *y,
                          double *acor, double *del);

/* ---- SPGMR on the device (replaces SUNLinSol_SPGMR + CVLS' difference-quotient Jv for this model;
 * reference settings src/Equations/cvode_config.cpp:172-179: PREC_NONE, maxl 5, modified Gram-Schmidt, no
 * restarts).  Solves (I - gamma J(t,y)) x = b scaled on both sides by ewt; J v = (f(y + sigma v) - fy)/sigma
 * through shud_b200_rhs_dev of `gpu`.  `tol` is on the 2-norm of the scaled residual (CVLS: eplifac * tq4 *
 * sqrt(N)).  One host synchronisation per Krylov iteration.  All vectors: device pointers, length ny(gpu).
 * returns 0 converged, 1 residual reduced but tol not met (SUNLS_RES_REDUCED), 2 not reduced, <0 error. ---- */
struct shud_ctx;
typedef struct shud_spgmr shud_spgmr;
int shud_spgmr_create(struct shud_ctx *gpu, shud_nvws *ws, int maxl, int64_t n_global, shud_spgmr **out);
void shud_spgmr_destroy(shud_spgmr *s);
/* global length of a distributed vector (sigma of the difference quotient is sqrt(N_global)); on a workspace with a
 * device allreduce (shud_nv_ws_set_allreduce) every dot product of the solve is reduced over the ranks on the device
 * and f() is shud_b200_rhs_exchange_dev */
void shud_spgmr_set_nglobal(shud_spgmr *s, int64_t n_global);
int shud_spgmr_solve(shud_spgmr *s, double t, double gamma, const double *y, const double *fy, const double *ewt,
                     const double *b, double tol, double *x, int *nli, double *resnorm);
/* One whole Newton iteration of the BDF corrector around y (fy = f(t, y)): the right-hand side
 * b = -(rl1 zn1 + acor - gamma fy) is formed, scaled and normed in one pass (never stored), solved as above, and the
 * correction applied in one pass: acor += x, y = zn0 + acor, *del = ||x||_WRMS(ewt) over n_global entries.  Element by
 * element the arithmetic of cvNlsResidual, N_VScale(-1), shud_spgmr_solve, N_VLinearSum x 2, N_VWrmsNorm.  Returns
 * shud_spgmr_solve's codes, or 3 when ||ewt b||_2 <= tol before the first iteration (nothing written). */
int shud_spgmr_newton_step(shud_spgmr *s, double t, double gamma, double rl1, const double *zn0, const double *zn1,
                           double *acor, double *y, const double *fy, const double *ewt, double tol, int64_t n_global,
                           double *del, int *nli, double *resnorm);
/* cvPredict (sgn = +1) / cvRestore (sgn = -1) of the Nordsieck array zn[0..q] in one pass, the in-place sums in the
 * order of the reference's N_VLinearSum calls; with acor != NULL also acor = 0, y = zn[0] + acor (start of cvNls). */
/* cvCompleteStep's zn[j] += l[j] acor (j = 0..q), the next step's weights ewt = 1 ./ (rtol |zn0| + atol), *nrm =
 * ||zn0||_WRMS(ewt) over n_global entries and, with yout != NULL, yout = zn0 - one pass. */
int shud_nv_bdf_complete(shud_nvws *ws, int64_t n, int q, const double *l, const double *acor, double *const *zn,
                         double rtol, double atol, double *ewt, double *yout, int64_t n_global, double *nrm);
int shud_nv_bdf_predict(shud_nvws *ws, int64_t n, int q, double sgn, double *const *zn, double *y, double *acor);

#ifdef __cplusplus
}
#endif
#endif
