/* TEST INFRASTRUCTURE ONLY (checker arm of the full-run comparisons; never loaded by the product).
 *
 * host_cv.c - what the reference's driver binds on the CPU, restated so the integrator of include/shud_cvode.h can
 * run the ORACLE arm with the very object code that runs the GPU arm:
 *   N_VNew_HostSerial : SUNDIALS' nvector_serial (N_VNew_Serial, src/Model/shud.cpp:59-64) - sequential loops in
 *                       index order, the definitions of the SUNDIALS 6 documentation (WrmsNorm = sqrt(sum((x w)^2)/N));
 *   shud_oracle_f     : the CVRhsFn int f(t, y, ydot, MD) (src/Model/f.hpp:12) on top of shud_oracle_rhs, with the
 *                       carried state (u_satn, qEleE_IC) living in the context as it lives in Model_Data.
 * SUNDIALS itself is absent from the reference tree and this image: "parity unpinned" for the vector arithmetic.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "shud_oracle.h"
#include "shud_sundials.h"
#include "shud_cvode.h"

typedef struct { sunindextype length; int own; double *data; long *opcount; } host_content;
#define HC(v) ((host_content *)(v)->content)
#define HD(v) (HC(v)->data)
#define HN(v) (HC(v)->length)

static long g_ops_count[64];
enum { OP_LINEARSUM, OP_CONST, OP_PROD, OP_DIV, OP_SCALE, OP_ABS, OP_INV, OP_ADDCONST, OP_DOT, OP_MAXNORM, OP_WRMS,
       OP_MIN, OP_LINCOMB, OP_SCALEADDMULTI, OP_DOTMULTI, OP_CLONE, OP_N };

static N_Vector host_new_empty(sunindextype n, SUNContext ctx);

static N_Vector_ID h_getid(N_Vector v) { (void)v; return SUNDIALS_NVEC_SERIAL; }
static N_Vector h_cloneempty(N_Vector w) {
    N_Vector v = host_new_empty(HN(w), w->sunctx);
    if (v) memcpy(v->ops, w->ops, sizeof(struct _generic_N_Vector_Ops));
    return v;
}
static N_Vector h_clone(N_Vector w) {
    N_Vector v = h_cloneempty(w);
    if (!v) return NULL;
    HC(v)->data = (double *)calloc((size_t)(HN(w) > 0 ? HN(w) : 1), sizeof(double));
    HC(v)->own = 1;
    g_ops_count[OP_CLONE]++;
    return v;
}
static void h_destroy(N_Vector v) {
    if (!v) return;
    if (HC(v)) { if (HC(v)->own) free(HC(v)->data); free(v->content); }
    free(v->ops);
    free(v);
}
static void h_space(N_Vector v, sunindextype *lrw, sunindextype *liw) { *lrw = HN(v); *liw = 1; }
static realtype *h_getarray(N_Vector v) { return HD(v); }
static void h_setarray(realtype *d, N_Vector v) { if (HC(v)->own) free(HD(v)); HC(v)->data = d; HC(v)->own = 0; }
static sunindextype h_getlength(N_Vector v) { return HN(v); }

static void h_linearsum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) {
    const sunindextype n = HN(z); const double *xd = HD(x), *yd = HD(y); double *zd = HD(z);
    g_ops_count[OP_LINEARSUM]++;
    for (sunindextype i = 0; i < n; i++) zd[i] = a * xd[i] + b * yd[i];
}
static void h_const(realtype c, N_Vector z) { g_ops_count[OP_CONST]++; for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = c; }
static void h_prod(N_Vector x, N_Vector y, N_Vector z) { g_ops_count[OP_PROD]++; for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = HD(x)[i] * HD(y)[i]; }
static void h_div(N_Vector x, N_Vector y, N_Vector z) { g_ops_count[OP_DIV]++; for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = HD(x)[i] / HD(y)[i]; }
static void h_scale(realtype c, N_Vector x, N_Vector z) {
    g_ops_count[OP_SCALE]++;
    if (z == x) { for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] *= c; return; }
    for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = c * HD(x)[i];
}
static void h_abs(N_Vector x, N_Vector z) { g_ops_count[OP_ABS]++; for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = fabs(HD(x)[i]); }
static void h_inv(N_Vector x, N_Vector z) { g_ops_count[OP_INV]++; for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = 1.0 / HD(x)[i]; }
static void h_addconst(N_Vector x, realtype b, N_Vector z) { g_ops_count[OP_ADDCONST]++; for (sunindextype i = 0; i < HN(z); i++) HD(z)[i] = HD(x)[i] + b; }
static realtype h_dot(N_Vector x, N_Vector y) {
    double s = 0.0; g_ops_count[OP_DOT]++;
    for (sunindextype i = 0; i < HN(x); i++) s += HD(x)[i] * HD(y)[i];
    return s;
}
static realtype h_maxnorm(N_Vector x) {
    double m = 0.0; g_ops_count[OP_MAXNORM]++;
    for (sunindextype i = 0; i < HN(x); i++) if (fabs(HD(x)[i]) > m) m = fabs(HD(x)[i]);
    return m;
}
static realtype h_wsqrsum(N_Vector x, N_Vector w) {
    double s = 0.0;
    for (sunindextype i = 0; i < HN(x); i++) { const double p = HD(x)[i] * HD(w)[i]; s += p * p; }
    return s;
}
static realtype h_wrms(N_Vector x, N_Vector w) { g_ops_count[OP_WRMS]++; return sqrt(h_wsqrsum(x, w) / (double)HN(x)); }
static realtype h_wl2(N_Vector x, N_Vector w) { return sqrt(h_wsqrsum(x, w)); }
static realtype h_l1(N_Vector x) { double s = 0.0; for (sunindextype i = 0; i < HN(x); i++) s += fabs(HD(x)[i]); return s; }
static realtype h_min(N_Vector x) {
    double m = DBL_MAX; g_ops_count[OP_MIN]++;
    for (sunindextype i = 0; i < HN(x); i++) if (HD(x)[i] < m) m = HD(x)[i];
    return m;
}
static int h_lincomb(int nvec, realtype *c, N_Vector *X, N_Vector z) {
    const sunindextype n = HN(z);
    g_ops_count[OP_LINCOMB]++;
    if (nvec < 1) return -1;
    /* nvector_serial: z = c0 X0, then z += ck Xk vector by vector (X0 may be z itself) */
    if (X[0] == z) { if (c[0] != 1.0) for (sunindextype i = 0; i < n; i++) HD(z)[i] *= c[0]; }
    else for (sunindextype i = 0; i < n; i++) HD(z)[i] = c[0] * HD(X[0])[i];
    for (int k = 1; k < nvec; k++) for (sunindextype i = 0; i < n; i++) HD(z)[i] += c[k] * HD(X[k])[i];
    return 0;
}
static int h_scaleaddmulti(int nvec, realtype *a, N_Vector x, N_Vector *Y, N_Vector *Z) {
    g_ops_count[OP_SCALEADDMULTI]++;
    for (int k = 0; k < nvec; k++) for (sunindextype i = 0; i < HN(x); i++) HD(Z[k])[i] = a[k] * HD(x)[i] + HD(Y[k])[i];
    return 0;
}
static int h_dotmulti(int nvec, N_Vector x, N_Vector *Y, realtype *d) {
    g_ops_count[OP_DOTMULTI]++;
    for (int k = 0; k < nvec; k++) { double s = 0.0; for (sunindextype i = 0; i < HN(x); i++) s += HD(x)[i] * HD(Y[k])[i]; d[k] = s; }
    return 0;
}

static N_Vector host_new_empty(sunindextype n, SUNContext ctx) {
    N_Vector v = (N_Vector)calloc(1, sizeof(struct _generic_N_Vector));
    if (!v) return NULL;
    v->ops = (N_Vector_Ops)calloc(1, sizeof(struct _generic_N_Vector_Ops));
    v->content = calloc(1, sizeof(host_content));
    if (!v->ops || !v->content) { free(v->ops); free(v->content); free(v); return NULL; }
    v->sunctx = ctx;
    HC(v)->length = n;
    N_Vector_Ops o = v->ops;
    o->nvgetvectorid = h_getid; o->nvclone = h_clone; o->nvcloneempty = h_cloneempty; o->nvdestroy = h_destroy;
    o->nvspace = h_space; o->nvgetarraypointer = h_getarray; o->nvsetarraypointer = h_setarray; o->nvgetlength = h_getlength;
    o->nvlinearsum = h_linearsum; o->nvconst = h_const; o->nvprod = h_prod; o->nvdiv = h_div; o->nvscale = h_scale;
    o->nvabs = h_abs; o->nvinv = h_inv; o->nvaddconst = h_addconst; o->nvdotprod = h_dot; o->nvmaxnorm = h_maxnorm;
    o->nvwrmsnorm = h_wrms; o->nvmin = h_min; o->nvwl2norm = h_wl2; o->nvl1norm = h_l1;
    o->nvlinearcombination = h_lincomb; o->nvscaleaddmulti = h_scaleaddmulti; o->nvdotprodmulti = h_dotmulti;
    o->nvdotprodlocal = h_dot; o->nvmaxnormlocal = h_maxnorm; o->nvminlocal = h_min; o->nvl1normlocal = h_l1;
    o->nvwsqrsumlocal = h_wsqrsum;
    return v;
}

N_Vector N_VNew_HostSerial(sunindextype n, SUNContext ctx) {
    N_Vector v = host_new_empty(n, ctx);
    if (!v) return NULL;
    HC(v)->data = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    HC(v)->own = 1;
    return v;
}
/* drop the fused members of one vector's table (CVODE then loops over the standard operations) */
void N_VDisableFused_HostSerial(N_Vector v) {
    v->ops->nvlinearcombination = NULL; v->ops->nvscaleaddmulti = NULL; v->ops->nvdotprodmulti = NULL;
}
long host_nv_opcount(int k) { return (k >= 0 && k < OP_N) ? g_ops_count[k] : -1; }

/* ---- the CVRhsFn of the oracle arm ---- */
typedef struct shud_oracle_model {
    const shud_mesh *mesh;
    const shud_forcing *forcing;
    double *u_satn, *qEleE_IC;   /* carried state [Ne] (Model_Data members in the reference) */
    int nthreads;
    long ncalls;
    int last_err;
} shud_oracle_model;

int shud_oracle_f(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
    shud_oracle_model *M = (shud_oracle_model *)user_data;
    (void)t;
    M->ncalls++;
    const int rc = shud_oracle_rhs(M->mesh, M->forcing, M->u_satn, M->qEleE_IC, N_VGetArrayPointer(y),
                                   N_VGetArrayPointer(ydot), NULL, M->nthreads);
    if (rc) { M->last_err = rc; return -1; }   /* the reference exits here (myexit); CVODE sees an unrecoverable failure */
    return 0;
}

/* a linear test problem for the integrator's own tests: ydot = -lambda .* y (+ optional coupling to the neighbour),
 * user_data = double[1 + n]: [kappa, lambda_0..lambda_n-1] */
int host_test_f(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
    const double *p = (const double *)user_data, kappa = p[0], *lam = p + 1;
    const double *yd = N_VGetArrayPointer(y);
    double *fd = N_VGetArrayPointer(ydot);
    const sunindextype n = N_VGetLength(y);
    (void)t;
    for (sunindextype i = 0; i < n; i++) {
        const double left = i > 0 ? yd[i - 1] : 0.0, right = i + 1 < n ? yd[i + 1] : 0.0;
        fd[i] = -lam[i] * yd[i] + kappa * (left - 2.0 * yd[i] + right) - 0.5 * yd[i] * yd[i] * yd[i];
    }
    return 0;
}

/* ---- The integrator's optional hooks (shud_cv_fused: predict, newton_step, ewt_set_norm - on the GPU single fused
 * kernels, shud_up_b200/csrc/shud_nvector_sundials.cu) restated through the generic vector operations, in the order
 * the unhooked integrator runs them.  Test infrastructure: tests/test_cvode_cpu.py checks that the hooked control
 * flow of shud_cvode.cpp reproduces the plain one bit for bit. ---- */
typedef struct host_fused_ctx { shud_cv *cv; N_Vector b, x; long calls[4]; } host_fused_ctx;

static int hf_predict(void *ctx, int q, realtype sgn, N_Vector *zn, N_Vector y, N_Vector acor) {
    ((host_fused_ctx *)ctx)->calls[0]++;
    for (int k = 1; k <= q; k++)
        for (int j = q; j >= k; j--) N_VLinearSum(1.0, zn[j - 1], sgn, zn[j], zn[j - 1]);
    if (acor) { N_VConst(0.0, acor); N_VLinearSum(1.0, zn[0], 1.0, acor, y); }
    return 0;
}
static int hf_newton_step(void *ctx, realtype t, realtype gamma, realtype rl1, N_Vector zn0, N_Vector zn1, N_Vector acor,
                          N_Vector y, N_Vector fy, N_Vector ewt, realtype delta, realtype *del, int *nli, int *nfe) {
    host_fused_ctx *c = (host_fused_ctx *)ctx;
    c->calls[1]++;
    N_VLinearSum(rl1, zn1, 1.0, acor, c->b);
    N_VLinearSum(-gamma, fy, 1.0, c->b, c->b);
    N_VScale(-1.0, c->b, c->b);
    int it = 0;
    const int r = shud_cv_linsolve(c->cv, t, gamma, y, fy, ewt, c->b, delta, c->x, &it);
    *nli = it; *nfe = 0;  /* the integrator's own difference quotient has counted its RHS calls */
    if (r < 0 || r == 3) return r;
    N_VLinearSum(1.0, acor, 1.0, c->x, acor);
    N_VLinearSum(1.0, zn0, 1.0, acor, y);
    *del = N_VWrmsNorm(c->x, ewt);
    return r;
}
static int hf_ewt_set_norm(void *ctx, realtype rtol, realtype atol, N_Vector y, N_Vector ewt, realtype *nrm) {
    host_fused_ctx *c = (host_fused_ctx *)ctx;
    c->calls[2]++;
    N_VAbs(y, c->b);
    N_VScale(rtol, c->b, c->b);
    N_VAddConst(c->b, atol, c->b);
    N_VInv(c->b, ewt);
    *nrm = N_VWrmsNorm(y, ewt);
    return 0;
}
static int hf_complete_step(void *ctx, int q, realtype *l, N_Vector acor, N_Vector *zn, realtype rtol, realtype atol,
                            N_Vector ewt_next, N_Vector yout, realtype *nrm) {
    host_fused_ctx *c = (host_fused_ctx *)ctx;
    c->calls[3]++;
    N_VScaleAddMulti(q + 1, l, acor, zn, zn);
    N_VAbs(zn[0], c->b);
    N_VScale(rtol, c->b, c->b);
    N_VAddConst(c->b, atol, c->b);
    N_VInv(c->b, ewt_next);
    *nrm = N_VWrmsNorm(zn[0], ewt_next);
    if (yout) N_VScale(1.0, zn[0], yout);
    return 0;
}
int host_cv_fused_create(shud_cv *cv, N_Vector tmpl, shud_cv_fused *out) {
    host_fused_ctx *c = (host_fused_ctx *)calloc(1, sizeof(host_fused_ctx));
    if (!c || !cv || !tmpl || !out) return -1;
    c->cv = cv; c->b = N_VClone(tmpl); c->x = N_VClone(tmpl);
    memset(out, 0, sizeof(*out));
    out->ctx = c; out->predict = hf_predict; out->newton_step = hf_newton_step; out->ewt_set_norm = hf_ewt_set_norm;
    out->complete_step = hf_complete_step;
    return 0;
}
void host_cv_fused_calls(const shud_cv_fused *f, long *calls4) {
    const host_fused_ctx *c = (const host_fused_ctx *)f->ctx;
    for (int k = 0; k < 4; k++) calls4[k] = c->calls[k];
}
void host_cv_fused_destroy(shud_cv_fused *f) {
    host_fused_ctx *c = (host_fused_ctx *)f->ctx;
    if (!c) return;
    N_VDestroy(c->b); N_VDestroy(c->x);
    free(c);
    f->ctx = NULL;
}
