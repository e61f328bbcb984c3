// shud_nvector_sundials.cu - the device N_Vector as a SUNDIALS 6 N_Vector: struct _generic_N_Vector_Ops filled with
// the flat shud_nv_* calls of shud_nvec.cu (include/shud_sundials.h; SURVEY.md 8(b) "N_Vector ops table").
//
// What binds here in the reference: N_VNew_Serial / N_VNew_OpenMP (src/Model/shud.cpp:59-64), the clones CVODE and
// SUNLinSol_SPGMR make from udata (src/Equations/cvode_config.cpp:169,176), and the host accesses NV_Ith_S /
// NV_DATA_S of SetIC2Y (src/ModelData/MD_initialize.cpp:117-135), summary (src/ModelData/MD_update.cpp:190-216)
// and the water-balance sampler (src/Model/shud.cpp:147).
//
// Host mirror: N_VGetArrayPointer returns a pinned host array in the reference's blocked order.  It is refreshed
// from the device only when a device operation has written the vector since the last refresh (every operation of the
// table marks its output), so element-wise host loops (NV_Ith_S) cost one transfer, not one per element.  Host writes
// through the pointer are pushed with N_VCopyToDevice_ShudB200 (the nvector_cuda convention).
// No CPU fallback: every arithmetic operation is a kernel of shud_nvec.cu.
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "shud_cvode.h"
#include "shud_sundials.h"

namespace {

struct Content : shud_nv_content {
    bool host_valid;  // the mirror holds what the device vector holds
};

long g_unused_calls = 0;
int g_check_rhs = 0;

inline Content *CT(N_Vector v) { return (Content *)v->content; }
inline double *D(N_Vector v) { return CT(v)->dev; }
inline shud_nvws *WS(N_Vector v) { return CT(v)->ws; }
inline int64_t LEN(N_Vector v) { return (int64_t)CT(v)->length; }
inline void wrote(N_Vector z) { CT(z)->host_valid = false; }

void fail(int rc, const char *what) {
    if (rc == 0) return;
    // SUNDIALS vector operations have no error channel: a failed launch is fatal, as a failed malloc is in nvector_serial
    fprintf(stderr, "[shud_b200] N_Vector operation %s failed (%d)\n", what, rc);
    abort();
}

double allreduce1(N_Vector v, double x, int op) {
    Content *c = CT(v);
    if (c->allreduce) fail(c->allreduce(c->comm, &x, 1, op), "allreduce");
    return x;
}

struct _generic_N_Vector_Ops g_ops;
bool g_ops_ready = false;
void fill_ops();

N_Vector new_empty(SUNContext sunctx) {
    if (!g_ops_ready) fill_ops();
    N_Vector v = (N_Vector)calloc(1, sizeof(struct _generic_N_Vector));
    if (!v) return nullptr;
    // every vector carries its own copy of the table, as N_VNewEmpty + N_VCopyOps do: CVODE may disable fused
    // operations per vector
    v->ops = (N_Vector_Ops)malloc(sizeof(struct _generic_N_Vector_Ops));
    if (!v->ops) { free(v); return nullptr; }
    memcpy(v->ops, &g_ops, sizeof(g_ops));
    v->sunctx = sunctx;
    Content *c = (Content *)calloc(1, sizeof(Content));
    if (!c) { free(v->ops); free(v); return nullptr; }
    v->content = c;
    return v;
}

// ---- constructors, destructors, utility operations ----
N_Vector_ID nv_getid(N_Vector) { return SUNDIALS_NVEC_CUSTOM; }

N_Vector nv_cloneempty(N_Vector w) {
    N_Vector v = new_empty(w->sunctx);
    if (!v) return nullptr;
    memcpy(v->ops, w->ops, sizeof(struct _generic_N_Vector_Ops));
    Content *c = CT(v), *cw = CT(w);
    c->length = cw->length; c->global_length = cw->global_length;
    c->ws = cw->ws; c->gpu = cw->gpu; c->allreduce = cw->allreduce; c->comm = cw->comm;
    c->own_dev = c->own_host = 0; c->dev = nullptr; c->host = nullptr; c->host_valid = false;
    return v;
}

void nv_destroy(N_Vector v) {
    if (!v) return;
    Content *c = CT(v);
    if (c) {
        if (c->ws) cudaSetDevice(shud_nv_ws_device(c->ws));
        if (c->own_dev && c->dev) {
            // work queued on the stream may still touch the buffer
            cudaStreamSynchronize((cudaStream_t)shud_nv_ws_stream(c->ws));
            cudaFree(c->dev);
        }
        if (c->own_host && c->host) cudaFreeHost(c->host);
        free(c);
    }
    free(v->ops);
    free(v);
}

N_Vector nv_clone(N_Vector w) {
    N_Vector v = nv_cloneempty(w);
    if (!v) return nullptr;
    Content *c = CT(v);
    cudaSetDevice(shud_nv_ws_device(c->ws));
    if (c->length > 0 && cudaMalloc(&c->dev, sizeof(double) * (size_t)c->length) != cudaSuccess) {
        cudaGetLastError();
        nv_destroy(v);
        return nullptr;
    }
    c->own_dev = 1;
    return v;
}

void nv_space(N_Vector v, sunindextype *lrw, sunindextype *liw) {
    if (lrw) *lrw = CT(v)->global_length;
    if (liw) *liw = 2;
}
sunindextype nv_getlength(N_Vector v) { return CT(v)->global_length; }
void *nv_getcommunicator(N_Vector v) { return CT(v)->comm; }
realtype *nv_getdevicearraypointer(N_Vector v) { return CT(v)->dev; }

int ensure_host(Content *c) {
    if (c->host) return 0;
    cudaSetDevice(shud_nv_ws_device(c->ws));
    if (cudaMallocHost(&c->host, sizeof(double) * (size_t)(c->length > 0 ? c->length : 1)) != cudaSuccess) return SHUD_ERR_CUDA;
    c->own_host = 1;
    c->host_valid = false;
    return 0;
}

realtype *nv_getarraypointer(N_Vector v) {
    Content *c = CT(v);
    if (!c->host_valid) fail(N_VCopyFromDevice_ShudB200(v), "N_VGetArrayPointer");
    return c->host;
}

void nv_setarraypointer(realtype *h, N_Vector v) {
    // attaches caller-owned HOST storage as the mirror (N_VSetArrayPointer of the serial vector replaces the data
    // array): the device copy is refreshed from it
    Content *c = CT(v);
    if (c->own_host && c->host) cudaFreeHost(c->host);
    c->host = h; c->own_host = 0; c->host_valid = false;
    if (h) fail(N_VCopyToDevice_ShudB200(v), "N_VSetArrayPointer");
}

// ---- standard vector operations ----
void nv_linearsum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) {
    fail(shud_nv_linearsum(WS(z), LEN(z), a, D(x), b, D(y), D(z)), "N_VLinearSum"); wrote(z);
}
void nv_const(realtype c, N_Vector z) { fail(shud_nv_const(WS(z), LEN(z), c, D(z)), "N_VConst"); wrote(z); }
void nv_prod(N_Vector x, N_Vector y, N_Vector z) { fail(shud_nv_prod(WS(z), LEN(z), D(x), D(y), D(z)), "N_VProd"); wrote(z); }
void nv_div(N_Vector x, N_Vector y, N_Vector z) { fail(shud_nv_div(WS(z), LEN(z), D(x), D(y), D(z)), "N_VDiv"); wrote(z); }
void nv_scale(realtype c, N_Vector x, N_Vector z) {
    if (z == x && c == 1.0) return;
    fail(shud_nv_scale(WS(z), LEN(z), c, D(x), D(z)), "N_VScale"); wrote(z);
}
void nv_abs(N_Vector x, N_Vector z) { fail(shud_nv_abs(WS(z), LEN(z), D(x), D(z)), "N_VAbs"); wrote(z); }
void nv_inv(N_Vector x, N_Vector z) { fail(shud_nv_inv(WS(z), LEN(z), D(x), D(z)), "N_VInv"); wrote(z); }
void nv_addconst(N_Vector x, realtype b, N_Vector z) { fail(shud_nv_addconst(WS(z), LEN(z), D(x), b, D(z)), "N_VAddConst"); wrote(z); }
void nv_compare(realtype c, N_Vector x, N_Vector z) {
    g_unused_calls++;
    fail(shud_nv_compare(WS(z), LEN(z), c, D(x), D(z)), "N_VCompare"); wrote(z);
}

// The flat reductions return the LOCAL value, or - when the workspace carries a device allreduce (a distributed vector
// on the library's own communicator, N_VSetDistributed_ShudB200) - the GLOBAL one.  LocalOnly brackets the *local
// members of the table; with a host allreduce hook (any other communicator) the global members are local + hook.
struct LocalOnly {
    shud_nvws *ws;
    explicit LocalOnly(shud_nvws *w) : ws(w) { shud_nv_ws_local(ws, 1); }
    ~LocalOnly() { shud_nv_ws_local(ws, 0); }
};
realtype f_dot(N_Vector x, N_Vector y) { double r; fail(shud_nv_dotprod(WS(x), LEN(x), D(x), D(y), &r), "N_VDotProd"); return r; }
realtype f_maxnorm(N_Vector x) { double r; fail(shud_nv_maxnorm(WS(x), LEN(x), D(x), &r), "N_VMaxNorm"); return r; }
realtype f_min(N_Vector x) { double r; fail(shud_nv_min(WS(x), LEN(x), D(x), &r), "N_VMin"); return r; }
realtype f_l1(N_Vector x) { double r; fail(shud_nv_l1norm(WS(x), LEN(x), D(x), &r), "N_VL1Norm"); return r; }
realtype f_wsqr(N_Vector x, N_Vector w) { double r; fail(shud_nv_wsqrsum(WS(x), LEN(x), D(x), D(w), &r), "N_VWSqrSumLocal"); return r; }
realtype f_wsqrmask(N_Vector x, N_Vector w, N_Vector id) {
    double r; fail(shud_nv_wsqrsum_mask(WS(x), LEN(x), D(x), D(w), D(id), &r), "N_VWSqrSumMaskLocal"); return r;
}
booleantype f_invtest(N_Vector x, N_Vector z) {
    int ok = 0; fail(shud_nv_invtest(WS(x), LEN(x), D(x), D(z), &ok), "N_VInvTest"); wrote(z); return ok ? SUNTRUE : SUNFALSE;
}
booleantype f_constrmask(N_Vector c, N_Vector x, N_Vector m) {
    int ok = 0; fail(shud_nv_constrmask(WS(x), LEN(x), D(c), D(x), D(m), &ok), "N_VConstrMask"); wrote(m); return ok ? SUNTRUE : SUNFALSE;
}
realtype f_minquot(N_Vector num, N_Vector den) {
    double r; fail(shud_nv_minquotient(WS(num), LEN(num), D(num), D(den), &r), "N_VMinQuotient"); return r;
}

// local reductions
realtype nv_dotprodlocal(N_Vector x, N_Vector y) { LocalOnly g(WS(x)); return f_dot(x, y); }
realtype nv_maxnormlocal(N_Vector x) { LocalOnly g(WS(x)); return f_maxnorm(x); }
realtype nv_minlocal(N_Vector x) { LocalOnly g(WS(x)); return f_min(x); }
realtype nv_l1normlocal(N_Vector x) { LocalOnly g(WS(x)); return f_l1(x); }
realtype nv_wsqrsumlocal(N_Vector x, N_Vector w) { LocalOnly g(WS(x)); return f_wsqr(x, w); }
realtype nv_wsqrsummasklocal(N_Vector x, N_Vector w, N_Vector id) { LocalOnly g(WS(x)); return f_wsqrmask(x, w, id); }
booleantype nv_invtestlocal(N_Vector x, N_Vector z) { LocalOnly g(WS(x)); return f_invtest(x, z); }
booleantype nv_constrmasklocal(N_Vector c, N_Vector x, N_Vector m) { LocalOnly g(WS(x)); return f_constrmask(c, x, m); }
realtype nv_minquotientlocal(N_Vector num, N_Vector den) { LocalOnly g(WS(num)); return f_minquot(num, den); }

// global reductions
realtype nv_dotprod(N_Vector x, N_Vector y) { return CT(x)->allreduce ? allreduce1(x, nv_dotprodlocal(x, y), 0) : f_dot(x, y); }
realtype nv_maxnorm(N_Vector x) { return CT(x)->allreduce ? allreduce1(x, nv_maxnormlocal(x), 1) : f_maxnorm(x); }
realtype nv_min(N_Vector x) { return CT(x)->allreduce ? allreduce1(x, nv_minlocal(x), 2) : f_min(x); }
realtype nv_l1norm(N_Vector x) { g_unused_calls++; return CT(x)->allreduce ? allreduce1(x, nv_l1normlocal(x), 0) : f_l1(x); }
realtype nv_wrmsnorm(N_Vector x, N_Vector w) {
    Content *c = CT(x);
    if (!c->allreduce) {  // the sqrt(sum / N_global) is done by the reduction kernel's last block (one GPU), or on the
                          // host behind the device allreduce (distributed vector on the library's communicator)
        double r; fail(shud_nv_wrmsnorm(c->ws, c->length, c->dev, D(w), c->global_length, &r), "N_VWrmsNorm"); return r;
    }
    return sqrt(allreduce1(x, nv_wsqrsumlocal(x, w), 0) / (double)c->global_length);
}
realtype nv_wrmsnormmask(N_Vector x, N_Vector w, N_Vector id) {
    g_unused_calls++;
    const double s = CT(x)->allreduce ? allreduce1(x, nv_wsqrsummasklocal(x, w, id), 0) : f_wsqrmask(x, w, id);
    return sqrt(s / (double)CT(x)->global_length);
}
realtype nv_wl2norm(N_Vector x, N_Vector w) {
    g_unused_calls++;
    return sqrt(CT(x)->allreduce ? allreduce1(x, nv_wsqrsumlocal(x, w), 0) : f_wsqr(x, w));
}
booleantype nv_invtest(N_Vector x, N_Vector z) {
    g_unused_calls++;
    if (!CT(x)->allreduce) return f_invtest(x, z);
    return allreduce1(x, nv_invtestlocal(x, z) ? 1.0 : 0.0, 2) > 0.5 ? SUNTRUE : SUNFALSE;
}
booleantype nv_constrmask(N_Vector c, N_Vector x, N_Vector m) {
    g_unused_calls++;
    if (!CT(x)->allreduce) return f_constrmask(c, x, m);
    return allreduce1(x, nv_constrmasklocal(c, x, m) ? 1.0 : 0.0, 2) > 0.5 ? SUNTRUE : SUNFALSE;
}
realtype nv_minquotient(N_Vector num, N_Vector den) {
    g_unused_calls++;
    return CT(num)->allreduce ? allreduce1(num, nv_minquotientlocal(num, den), 2) : f_minquot(num, den);
}

// ---- fused and vector-array operations (nvec <= SHUD_NV_MAXVEC per launch; longer lists in chunks) ----
int nv_linearcombination(int nvec, realtype *c, N_Vector *X, N_Vector z) {
    if (nvec < 1) return -1;
    const double *p[SHUD_NV_MAXVEC];
    if (nvec <= SHUD_NV_MAXVEC) {
        for (int k = 0; k < nvec; k++) p[k] = D(X[k]);
        fail(shud_nv_linearcombination(WS(z), LEN(z), nvec, c, p, D(z)), "N_VLinearCombination");
    } else {
        // z = sum of the first chunk, then z += the following chunks (z itself rides as vector 0 with coefficient 1)
        int k0 = 0;
        bool first = true;
        while (k0 < nvec) {
            double cc[SHUD_NV_MAXVEC];
            int n = 0;
            if (!first) { cc[0] = 1.0; p[0] = D(z); n = 1; }
            while (n < SHUD_NV_MAXVEC && k0 < nvec) { cc[n] = c[k0]; p[n] = D(X[k0]); n++; k0++; }
            fail(shud_nv_linearcombination(WS(z), LEN(z), n, cc, p, D(z)), "N_VLinearCombination");
            first = false;
        }
    }
    wrote(z);
    return 0;
}
int nv_scaleaddmulti(int nvec, realtype *a, N_Vector x, N_Vector *Y, N_Vector *Z) {
    for (int k0 = 0; k0 < nvec; k0 += SHUD_NV_MAXVEC) {
        const int n = nvec - k0 < SHUD_NV_MAXVEC ? nvec - k0 : SHUD_NV_MAXVEC;
        const double *py[SHUD_NV_MAXVEC]; double *pz[SHUD_NV_MAXVEC];
        for (int k = 0; k < n; k++) { py[k] = D(Y[k0 + k]); pz[k] = D(Z[k0 + k]); wrote(Z[k0 + k]); }
        fail(shud_nv_scaleaddmulti(WS(x), LEN(x), n, a + k0, D(x), py, pz), "N_VScaleAddMulti");
    }
    return 0;
}
int dotprodmulti_flat(int nvec, N_Vector x, N_Vector *Y, realtype *out) {
    for (int k0 = 0; k0 < nvec; k0 += SHUD_NV_MAXVEC) {
        const int n = nvec - k0 < SHUD_NV_MAXVEC ? nvec - k0 : SHUD_NV_MAXVEC;
        const double *py[SHUD_NV_MAXVEC];
        for (int k = 0; k < n; k++) py[k] = D(Y[k0 + k]);
        fail(shud_nv_dotprodmulti(WS(x), LEN(x), n, D(x), py, out + k0), "N_VDotProdMulti");
    }
    return 0;
}
int nv_dotprodmultilocal(int nvec, N_Vector x, N_Vector *Y, realtype *out) {
    LocalOnly g(WS(x));
    return dotprodmulti_flat(nvec, x, Y, out);
}
int nv_dotprodmultiallreduce(int nvec, N_Vector x, realtype *sum) {
    Content *c = CT(x);
    if (c->allreduce) fail(c->allreduce(c->comm, sum, nvec, 0), "allreduce");
    return 0;
}
int nv_dotprodmulti(int nvec, N_Vector x, N_Vector *Y, realtype *out) {
    // ONE allreduce for all the dot products of a Gram-Schmidt sweep
    if (!CT(x)->allreduce) return dotprodmulti_flat(nvec, x, Y, out);  // global already behind a device allreduce
    nv_dotprodmultilocal(nvec, x, Y, out);
    return nv_dotprodmultiallreduce(nvec, x, out);
}
int nv_linearsumvectorarray(int nvec, realtype a, N_Vector *X, realtype b, N_Vector *Y, N_Vector *Z) {
    for (int k0 = 0; k0 < nvec; k0 += SHUD_NV_MAXVEC) {
        const int n = nvec - k0 < SHUD_NV_MAXVEC ? nvec - k0 : SHUD_NV_MAXVEC;
        const double *px[SHUD_NV_MAXVEC], *py[SHUD_NV_MAXVEC]; double *pz[SHUD_NV_MAXVEC];
        for (int k = 0; k < n; k++) { px[k] = D(X[k0 + k]); py[k] = D(Y[k0 + k]); pz[k] = D(Z[k0 + k]); wrote(Z[k0 + k]); }
        fail(shud_nv_linearsumvectorarray(WS(Z[0]), LEN(Z[0]), n, a, px, b, py, pz), "N_VLinearSumVectorArray");
    }
    return 0;
}
int nv_scalevectorarray(int nvec, realtype *c, N_Vector *X, N_Vector *Z) {
    for (int k0 = 0; k0 < nvec; k0 += SHUD_NV_MAXVEC) {
        const int n = nvec - k0 < SHUD_NV_MAXVEC ? nvec - k0 : SHUD_NV_MAXVEC;
        const double *px[SHUD_NV_MAXVEC]; double *pz[SHUD_NV_MAXVEC];
        for (int k = 0; k < n; k++) { px[k] = D(X[k0 + k]); pz[k] = D(Z[k0 + k]); wrote(Z[k0 + k]); }
        fail(shud_nv_scalevectorarray(WS(Z[0]), LEN(Z[0]), n, c + k0, px, pz), "N_VScaleVectorArray");
    }
    return 0;
}
int nv_constvectorarray(int nvec, realtype c, N_Vector *Z) {
    for (int k0 = 0; k0 < nvec; k0 += SHUD_NV_MAXVEC) {
        const int n = nvec - k0 < SHUD_NV_MAXVEC ? nvec - k0 : SHUD_NV_MAXVEC;
        double *pz[SHUD_NV_MAXVEC];
        for (int k = 0; k < n; k++) { pz[k] = D(Z[k0 + k]); wrote(Z[k0 + k]); }
        fail(shud_nv_constvectorarray(WS(Z[0]), LEN(Z[0]), n, c, pz), "N_VConstVectorArray");
    }
    return 0;
}
int nv_wrmsnormvectorarray(int nvec, N_Vector *X, N_Vector *W, realtype *out) {
    Content *c = CT(X[0]);
    for (int k0 = 0; k0 < nvec; k0 += SHUD_NV_MAXVEC) {
        const int n = nvec - k0 < SHUD_NV_MAXVEC ? nvec - k0 : SHUD_NV_MAXVEC;
        const double *px[SHUD_NV_MAXVEC], *pw[SHUD_NV_MAXVEC];
        for (int k = 0; k < n; k++) { px[k] = D(X[k0 + k]); pw[k] = D(W[k0 + k]); }
        if (!c->allreduce) {
            fail(shud_nv_wrmsnormvectorarray(c->ws, c->length, n, px, pw, c->global_length, out + k0), "N_VWrmsNormVectorArray");
        } else {
            for (int k = 0; k < n; k++) out[k0 + k] = nv_wsqrsumlocal(X[k0 + k], W[k0 + k]);
        }
    }
    if (c->allreduce) {
        fail(c->allreduce(c->comm, out, nvec, 0), "allreduce");
        for (int k = 0; k < nvec; k++) out[k] = sqrt(out[k] / (double)c->global_length);
    }
    return 0;
}

void nv_print(N_Vector v) {
    const double *h = nv_getarraypointer(v);
    for (int64_t i = 0; i < LEN(v); i++) printf("%.16g\n", h[i]);
}
void nv_printfile(N_Vector v, FILE *f) {
    const double *h = nv_getarraypointer(v);
    for (int64_t i = 0; i < LEN(v); i++) fprintf(f, "%.16g\n", h[i]);
}

void fill_ops() {
    memset(&g_ops, 0, sizeof(g_ops));
    g_ops.nvgetvectorid = nv_getid;
    g_ops.nvclone = nv_clone; g_ops.nvcloneempty = nv_cloneempty; g_ops.nvdestroy = nv_destroy; g_ops.nvspace = nv_space;
    g_ops.nvgetarraypointer = nv_getarraypointer; g_ops.nvgetdevicearraypointer = nv_getdevicearraypointer;
    g_ops.nvsetarraypointer = nv_setarraypointer; g_ops.nvgetcommunicator = nv_getcommunicator;
    g_ops.nvgetlength = nv_getlength;
    g_ops.nvlinearsum = nv_linearsum; g_ops.nvconst = nv_const; g_ops.nvprod = nv_prod; g_ops.nvdiv = nv_div;
    g_ops.nvscale = nv_scale; g_ops.nvabs = nv_abs; g_ops.nvinv = nv_inv; g_ops.nvaddconst = nv_addconst;
    g_ops.nvdotprod = nv_dotprod; g_ops.nvmaxnorm = nv_maxnorm; g_ops.nvwrmsnorm = nv_wrmsnorm;
    g_ops.nvwrmsnormmask = nv_wrmsnormmask; g_ops.nvmin = nv_min; g_ops.nvwl2norm = nv_wl2norm; g_ops.nvl1norm = nv_l1norm;
    g_ops.nvcompare = nv_compare; g_ops.nvinvtest = nv_invtest; g_ops.nvconstrmask = nv_constrmask;
    g_ops.nvminquotient = nv_minquotient;
    g_ops.nvlinearcombination = nv_linearcombination; g_ops.nvscaleaddmulti = nv_scaleaddmulti;
    g_ops.nvdotprodmulti = nv_dotprodmulti;
    g_ops.nvlinearsumvectorarray = nv_linearsumvectorarray; g_ops.nvscalevectorarray = nv_scalevectorarray;
    g_ops.nvconstvectorarray = nv_constvectorarray; g_ops.nvwrmsnormvectorarray = nv_wrmsnormvectorarray;
    // nvwrmsnormmaskvectorarray, nvscaleaddmultivectorarray, nvlinearcombinationvectorarray stay NULL: SUNDIALS then
    // falls back to loops over the operations above (they are sensitivity-analysis operations CVODE does not call)
    g_ops.nvdotprodlocal = nv_dotprodlocal; g_ops.nvmaxnormlocal = nv_maxnormlocal; g_ops.nvminlocal = nv_minlocal;
    g_ops.nvl1normlocal = nv_l1normlocal; g_ops.nvinvtestlocal = nv_invtestlocal;
    g_ops.nvconstrmasklocal = nv_constrmasklocal; g_ops.nvminquotientlocal = nv_minquotientlocal;
    g_ops.nvwsqrsumlocal = nv_wsqrsumlocal; g_ops.nvwsqrsummasklocal = nv_wsqrsummasklocal;
    g_ops.nvdotprodmultilocal = nv_dotprodmultilocal; g_ops.nvdotprodmultiallreduce = nv_dotprodmultiallreduce;
    g_ops.nvprint = nv_print; g_ops.nvprintfile = nv_printfile;
    g_ops_ready = true;
}

}  // namespace

extern "C" {

N_Vector N_VMake_ShudB200(sunindextype length, double *dev, shud_nvws *ws, struct shud_ctx *gpu, SUNContext sunctx) {
    if (length < 0 || !ws) return nullptr;
    if (gpu && (int64_t)length != shud_b200_ny(gpu)) return nullptr;
    N_Vector v = new_empty(sunctx);
    if (!v) return nullptr;
    Content *c = CT(v);
    c->length = c->global_length = length;
    c->ws = ws; c->gpu = gpu; c->dev = dev;
    return v;
}

N_Vector N_VNew_ShudB200(sunindextype length, shud_nvws *ws, struct shud_ctx *gpu, SUNContext sunctx) {
    N_Vector v = N_VMake_ShudB200(length, nullptr, ws, gpu, sunctx);
    if (!v) return nullptr;
    Content *c = CT(v);
    cudaSetDevice(shud_nv_ws_device(ws));
    if (length > 0 && cudaMalloc(&c->dev, sizeof(double) * (size_t)length) != cudaSuccess) {
        cudaGetLastError();
        nv_destroy(v);
        return nullptr;
    }
    c->own_dev = 1;
    if (length > 0) cudaMemsetAsync(c->dev, 0, sizeof(double) * (size_t)length, (cudaStream_t)shud_nv_ws_stream(ws));
    return v;
}

void N_VSetDistributed_ShudB200(N_Vector v, sunindextype global_length, shud_nv_allreduce_fn fn, void *comm) {
    Content *c = CT(v);
    c->global_length = global_length; c->comm = comm;
    if (fn == shud_b200_nv_allreduce) {
        // the library's own communicator: partial results never leave the device before they are reduced
        // (ncclAllReduce on the vectors' stream, one host synchronisation per reduction)
        c->allreduce = nullptr;
        shud_nv_ws_set_allreduce(c->ws, shud_b200_allreduce_dev, comm);
        // ... and, when the ranks' peer-to-peer blocks are mapped (shud_b200_p2p_connect), not even a second kernel:
        // the reduction kernels combine the ranks' partials themselves over NVLink (SHUD_P2P_AR=0: NCCL as above)
        int nr = 0, rk = 0;
        void *boxes[SHUD_NV_MAXRANKS] = {nullptr};
        const char *env = getenv("SHUD_P2P_AR");
        if (!(env && atoi(env) == 0) && shud_b200_p2p_mailboxes((shud_ctx *)comm, &nr, &rk, boxes) == 0 && nr > 1)
            shud_nv_ws_set_peer_allreduce(c->ws, nr, rk, boxes);
        else
            shud_nv_ws_set_peer_allreduce(c->ws, 0, 0, nullptr);
    } else {
        c->allreduce = fn;
    }
}

int N_VCopyToDevice_ShudB200(N_Vector v) {
    Content *c = CT(v);
    if (!c->host) return SHUD_ERR_ARG;
    if (c->length == 0) return 0;
    cudaSetDevice(shud_nv_ws_device(c->ws));
    int rc = 0;
    if (c->gpu) {
        rc = shud_b200_upload_ref(c->gpu, c->host, c->dev);
    } else {
        cudaStream_t st = (cudaStream_t)shud_nv_ws_stream(c->ws);
        if (cudaMemcpyAsync(c->dev, c->host, sizeof(double) * (size_t)c->length, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) rc = SHUD_ERR_CUDA;
    }
    if (rc == 0) c->host_valid = true;
    return rc;
}

int N_VCopyFromDevice_ShudB200(N_Vector v) {
    Content *c = CT(v);
    int rc = ensure_host(c);
    if (rc) return rc;
    if (c->length == 0) { c->host_valid = true; return 0; }
    cudaSetDevice(shud_nv_ws_device(c->ws));
    if (c->gpu) {
        rc = shud_b200_download_ref(c->gpu, c->dev, c->host);
    } else {
        cudaStream_t st = (cudaStream_t)shud_nv_ws_stream(c->ws);
        if (cudaMemcpyAsync(c->host, c->dev, sizeof(double) * (size_t)c->length, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) rc = SHUD_ERR_CUDA;
    }
    if (rc == 0) c->host_valid = true;
    return rc;
}

double *N_VSummary_ShudB200(N_Vector v) {
    Content *c = CT(v);
    if (!c->gpu || ensure_host(c)) return nullptr;
    if (shud_b200_summary_dev(c->gpu, c->dev, c->host)) return nullptr;
    c->host_valid = false;  // the mirror now differs from the device vector where a BC value was substituted
    return c->host;
}

double *N_VGetDeviceArrayPointer_ShudB200(N_Vector v) { return v ? CT(v)->dev : nullptr; }
long N_VOpsCalled_ShudB200(void) { return g_unused_calls; }

void shud_b200_f_check(int on) { g_check_rhs = on; }

int shud_b200_f(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
    shud_ctx *gpu = (shud_ctx *)user_data;
    if (!gpu || !y || !ydot) return -1;
    if (shud_b200_rhs_dev(gpu, t, CT(y)->dev, CT(ydot)->dev) != 0) return -1;
    wrote(ydot);
    if (g_check_rhs) {
        int32_t where = 0;
        if (shud_b200_check(gpu, &where) != 0) return -1;
    }
    return 0;
}

// the same on a partition of a multi-GPU run: halo exchange + RHS (shud_b200_rhs_exchange_dev)
int shud_b200_f_exchange(realtype t, N_Vector y, N_Vector ydot, void *user_data) {
    shud_ctx *gpu = (shud_ctx *)user_data;
    if (!gpu || !y || !ydot) return -1;
    if (shud_b200_rhs_exchange_dev(gpu, t, CT(y)->dev, CT(ydot)->dev) != 0) return -1;
    wrote(ydot);
    if (g_check_rhs) {
        int32_t where = 0;
        if (shud_b200_check(gpu, &where) != 0) return -1;
    }
    return 0;
}

// ---- fused operations of the integrator on the device (shud_cv_fused, include/shud_cvode.h) ----
struct cv_fused_ctx { shud_ctx *gpu; shud_nvws *ws; shud_spgmr *spgmr; };

static int fused_ewt_set(void *ctx, realtype rtol, realtype atol, N_Vector y, N_Vector ewt) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    const int rc = shud_nv_ewt(c->ws, LEN(y), rtol, atol, D(y), D(ewt));
    wrote(ewt);
    return rc;
}
static int fused_ewt_set_norm(void *ctx, realtype rtol, realtype atol, N_Vector y, N_Vector ewt, realtype *nrm) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    const int rc = shud_nv_ewt_wrms(c->ws, LEN(y), rtol, atol, D(y), D(ewt), CT(y)->global_length, nrm);
    wrote(ewt);
    return rc;
}
static int fused_complete_step(void *ctx, int q, realtype *l, N_Vector acor, N_Vector *zn, realtype rtol, realtype atol,
                               N_Vector ewt_next, N_Vector yout, realtype *nrm) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    if (q < 1 || q > 5) return -1;
    double *z[6];
    for (int j = 0; j <= q; j++) z[j] = D(zn[j]);
    const int rc = shud_nv_bdf_complete(c->ws, LEN(acor), q, l, D(acor), z, rtol, atol, D(ewt_next), yout ? D(yout) : nullptr,
                                        CT(acor)->global_length, nrm);
    for (int j = 0; j <= q; j++) wrote(zn[j]);
    wrote(ewt_next);
    if (yout) wrote(yout);
    return rc;
}
static int fused_nls_residual(void *ctx, realtype rl1, N_Vector zn1, N_Vector ycor, realtype gamma, N_Vector f, N_Vector res) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    // (rl1 zn1 + ycor) + (-gamma f): the two N_VLinearSum calls of cvNlsResidual in one pass, same operation order
    const double cf[3] = {rl1, 1.0, -gamma};
    const double *X[3] = {D(zn1), D(ycor), D(f)};
    const int rc = shud_nv_linearcombination(c->ws, LEN(res), 3, cf, X, D(res));
    wrote(res);
    return rc;
}
static int fused_lsolve(void *ctx, realtype t, realtype gamma, N_Vector y, N_Vector fy, N_Vector ewt, N_Vector b,
                        realtype delta, N_Vector x, int *nli, int *nfe) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    double res = 0.0;
    int it = 0;
    shud_spgmr_set_nglobal(c->spgmr, CT(y)->global_length);
    const int rc = shud_spgmr_solve(c->spgmr, t, gamma, D(y), D(fy), D(ewt), D(b), delta, D(x), &it, &res);
    wrote(x);
    if (nli) *nli = it;
    if (nfe) *nfe = it;
    if (rc == 0 && it == 0) return 3;  // ||s b||_2 <= delta before the first iteration (cvLsSolve's norm test)
    return rc;
}

static int fused_predict(void *ctx, int q, realtype sgn, N_Vector *zn, N_Vector y, N_Vector acor) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    if (q < 1 || q > 5) return -1;
    double *z[6];
    for (int j = 0; j <= q; j++) z[j] = D(zn[j]);
    const int rc = shud_nv_bdf_predict(c->ws, LEN(zn[0]), q, sgn, z, y ? D(y) : nullptr, acor ? D(acor) : nullptr);
    for (int j = 0; j < q; j++) wrote(zn[j]);
    if (acor) { wrote(acor); wrote(y); }
    return rc;
}
static int fused_newton_step(void *ctx, realtype t, realtype gamma, realtype rl1, N_Vector zn0, N_Vector zn1, N_Vector acor,
                             N_Vector y, N_Vector fy, N_Vector ewt, realtype delta, realtype *del, int *nli, int *nfe) {
    cv_fused_ctx *c = (cv_fused_ctx *)ctx;
    double res = 0.0;
    int it = 0;
    shud_spgmr_set_nglobal(c->spgmr, CT(y)->global_length);
    const int rc = shud_spgmr_newton_step(c->spgmr, t, gamma, rl1, D(zn0), D(zn1), D(acor), D(y), D(fy), D(ewt), delta,
                                          CT(y)->global_length, del, &it, &res);
    if (nli) *nli = it;
    if (nfe) *nfe = it;
    if (rc != 3 && rc >= 0) { wrote(acor); wrote(y); }
    return rc;
}

int shud_b200_cv_fused_create(shud_ctx *gpu, shud_nvws *ws, int maxl, shud_cv_fused *out) {
    if (!gpu || !ws || !out) return SHUD_ERR_ARG;
    cv_fused_ctx *c = (cv_fused_ctx *)calloc(1, sizeof(cv_fused_ctx));
    if (!c) return SHUD_ERR_ARG;
    c->gpu = gpu; c->ws = ws;
    const int rc = shud_spgmr_create(gpu, ws, maxl > 0 ? maxl : 5, shud_b200_ny(gpu), &c->spgmr);
    if (rc) { free(c); return rc; }
    out->ctx = c; out->ewt_set = fused_ewt_set; out->nls_residual = fused_nls_residual; out->lsolve = fused_lsolve;
    out->predict = fused_predict; out->newton_step = fused_newton_step; out->ewt_set_norm = fused_ewt_set_norm;
    out->complete_step = fused_complete_step;
    return SHUD_OK;
}
void shud_b200_cv_fused_destroy(shud_cv_fused *f) {
    if (!f || !f->ctx) return;
    cv_fused_ctx *c = (cv_fused_ctx *)f->ctx;
    shud_spgmr_destroy(c->spgmr);
    free(c);
    f->ctx = nullptr;
}

// Allreduce of a few host doubles over the context's NCCL communicator: the hook of a distributed vector
// (N_VSetDistributed_ShudB200(v, n_global, shud_b200_nv_allreduce, ctx))
int shud_b200_nv_allreduce(void *comm, double *vals, int n, int op) {
    return shud_b200_allreduce((shud_ctx *)comm, vals, n, op);
}

}  // extern "C"
