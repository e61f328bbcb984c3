/* cvode_from_c.c - the SUNDIALS-facing side of the boundary used from plain C, in the order the reference's driver
 * binds it (src/Model/shud.cpp:59-68,78,131,137):
 *   N_VNew_Serial(NY)            -> N_VNew_ShudB200(NY, ws, gpu)
 *   SetIC2Y (NV_Ith_S writes)    -> N_VGetArrayPointer + N_VCopyToDevice_ShudB200
 *   CVODE / SPGMR clones         -> v->ops->nvclone
 *   every vector operation       -> v->ops->nv...  (checked here against the flat shud_nv_* calls, bit for bit)
 *   f(t, y, ydot, MD)            -> shud_b200_f(t, y, ydot, gpu)
 *   CVode(mem, tout, ...)        -> shud_cv_solve (include/shud_cvode.h), plain ops-table path and device-fused path
 *   summary(udata)               -> N_VSummary_ShudB200
 *
 *   gcc -O2 -I include examples/cvode_from_c.c -L shud_up_b200 -lshud_b200 -Wl,-rpath,$PWD/shud_up_b200 -lm -o cvode_from_c
 *   ./cvode_from_c mesh.shudb200 case.bin [minutes]
 * case.bin as for rhs_from_c.c.  Exit code 3 = no CUDA device (there is no CPU path); 4 = a check failed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "shud_cvode.h"

static double *read_doubles(FILE *fp, size_t n) {
    double *p = (double *)malloc(sizeof(double) * (n ? n : 1));
    if (!p || fread(p, sizeof(double), n, fp) != n) { free(p); return NULL; }
    return p;
}
static int fails = 0;
static void check(int ok, const char *what) {
    printf("%s %s\n", ok ? "ok  " : "FAIL", what);
    if (!ok) fails++;
}
/* device buffer -> malloc'ed host copy, through a throw-away vector that wraps it (no context: no permutation) */
static double *fetch(double *dev, size_t n, shud_nvws *ws) {
    N_Vector w = N_VMake_ShudB200((sunindextype)n, dev, ws, NULL, NULL);
    double *h = (double *)malloc(sizeof(double) * n);
    memcpy(h, N_VGetArrayPointer(w), sizeof(double) * n);
    N_VDestroy(w);
    return h;
}
static int same(const double *a, const double *b, size_t n) { return memcmp(a, b, sizeof(double) * n) == 0; }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s mesh.shudb200 case.bin [minutes]\n", argv[0]); return 2; }
    const double t_end = argc > 3 ? atof(argv[3]) : 30.0;
    shud_mesh M;
    void *block = NULL;
    if (shud_b200_mesh_load(argv[1], &M, &block) != SHUD_OK) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const size_t Ne = (size_t)M.Ne, NY = 3 * Ne + (size_t)M.Nr + (size_t)M.Nl;
    FILE *fp = fopen(argv[2], "rb");
    if (!fp) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    double *y0 = read_doubles(fp, NY), *arr[8];
    int ok = y0 != NULL;
    for (int k = 0; k < 8; k++) { arr[k] = ok ? read_doubles(fp, Ne) : NULL; ok = ok && arr[k] != NULL; }
    fclose(fp);
    if (!ok) { fprintf(stderr, "short case file\n"); return 2; }

    shud_ctx *gpu = NULL;
    int rc = shud_b200_create(&M, 0, &gpu);
    if (rc == SHUD_ERR_NO_DEVICE) { fprintf(stderr, "no CUDA device: the library has no CPU fallback\n"); return 3; }
    if (rc != SHUD_OK) { fprintf(stderr, "shud_b200_create failed: %d\n", rc); return 1; }
    shud_forcing F = {0};
    F.qEleNetPrep = arr[0]; F.qPotEvap = arr[1]; F.qPotTran = arr[2]; F.t_lai = arr[3]; F.fu_Surf = arr[4]; F.fu_Sub = arr[5];
    F.qElePrep = arr[6]; F.qEleE_IC = arr[7];
    if (shud_b200_set_forcing(gpu, &F) != SHUD_OK || shud_b200_prime(gpu, y0) != SHUD_OK) return 1;

    /* ---- vectors: N_VNew, SetIC2Y through the host mirror, clones ---- */
    shud_nvws *ws = NULL;
    if (shud_nv_ws_create(0, shud_b200_stream(gpu), &ws) != SHUD_OK) return 1;
    N_Vector udata = N_VNew_ShudB200((sunindextype)NY, ws, gpu, NULL);
    if (!udata) return 1;
    double *mirror = N_VGetArrayPointer(udata);
    for (size_t i = 0; i < NY; i++) mirror[i] = y0[i];                   /* NV_Ith_S(udata, i) = ... */
    if (N_VCopyToDevice_ShudB200(udata) != 0) return 1;
    N_Vector du = N_VClone(udata), z = N_VClone(udata), w = N_VClone(udata);
    check(du && z && w, "nvclone");
    check(N_VGetLength(udata) == (sunindextype)NY && udata->ops->nvgetvectorid(udata) == SUNDIALS_NVEC_CUSTOM, "nvgetlength / nvgetvectorid");
    sunindextype lrw = 0, liw = 0;
    udata->ops->nvspace(udata, &lrw, &liw);
    check(lrw == (sunindextype)NY, "nvspace");
    N_VScale(1.0, udata, z);                                              /* device copy, then read back through the mirror */
    check(same(N_VGetArrayPointer(z), y0, NY), "host mirror round trip (reference order in, reference order out)");

    /* ---- f(): CVRhsFn on the vectors ---- */
    shud_b200_f_check(1);
    check(shud_b200_f(0.0, udata, du, gpu) == 0, "shud_b200_f");
    {
        const double *d = N_VGetArrayPointer(du);
        double s = 0, sa = 0;
        for (size_t i = 0; i < NY; i++) { s += d[i]; sa += fabs(d[i]); }
        printf("f: sum(ydot)=%.17g sum|ydot|=%.17g\n", s, sa);
    }
    shud_b200_prime(gpu, y0);  /* carried state back to where the comparison run starts */

    /* ---- the operations table against the flat calls, bit for bit ---- */
    double *dz = N_VGetDeviceArrayPointer_ShudB200(z), *dw = N_VGetDeviceArrayPointer_ShudB200(w);
    double *dU = N_VGetDeviceArrayPointer_ShudB200(udata), *dD = N_VGetDeviceArrayPointer_ShudB200(du);
    udata->ops->nvlinearsum(0.7, udata, -1.3, du, z);
    shud_nv_linearsum(ws, (int64_t)NY, 0.7, dU, -1.3, dD, dw);
    { double *a = fetch(dz, NY, ws), *b = fetch(dw, NY, ws); check(same(a, b, NY), "nvlinearsum == shud_nv_linearsum"); free(a); free(b); }
    udata->ops->nvabs(udata, z); udata->ops->nvaddconst(z, 1e-4, z); udata->ops->nvinv(z, z);   /* an ewt-like weight in z */
    shud_nv_abs(ws, (int64_t)NY, dU, dw); shud_nv_addconst(ws, (int64_t)NY, dw, 1e-4, dw); shud_nv_inv(ws, (int64_t)NY, dw, dw);
    { double *a = fetch(dz, NY, ws), *b = fetch(dw, NY, ws); check(same(a, b, NY), "nvabs / nvaddconst / nvinv == flat"); free(a); free(b); }
    double r1, r2;
    r1 = udata->ops->nvwrmsnorm(du, z); shud_nv_wrmsnorm(ws, (int64_t)NY, dD, dz, (int64_t)NY, &r2);
    check(r1 == r2 && r1 > 0, "nvwrmsnorm == shud_nv_wrmsnorm");
    r1 = udata->ops->nvdotprod(udata, du); shud_nv_dotprod(ws, (int64_t)NY, dU, dD, &r2);
    check(r1 == r2, "nvdotprod == shud_nv_dotprod");
    r1 = udata->ops->nvmaxnorm(du); shud_nv_maxnorm(ws, (int64_t)NY, dD, &r2);
    check(r1 == r2, "nvmaxnorm == shud_nv_maxnorm");
    r1 = udata->ops->nvmin(udata); shud_nv_min(ws, (int64_t)NY, dU, &r2);
    check(r1 == r2, "nvmin == shud_nv_min");
    {
        /* host loops over the mirror pin the definitions (SUNDIALS: WrmsNorm = sqrt(sum((x w)^2) / N)) */
        const double *hd = N_VGetArrayPointer(du), *hz = N_VGetArrayPointer(z);
        double s = 0, mx = 0;
        for (size_t i = 0; i < NY; i++) { s += hd[i] * hz[i] * hd[i] * hz[i]; if (fabs(hd[i]) > mx) mx = fabs(hd[i]); }
        const double wr = udata->ops->nvwrmsnorm(du, z);
        check(fabs(wr - sqrt(s / (double)NY)) <= 1e-13 * wr && udata->ops->nvmaxnorm(du) == mx, "nvwrmsnorm / nvmaxnorm == host definition");
    }
    {
        double c[3] = {0.5, -2.0, 3.0}, d3[3], e3[3];
        N_Vector X[3] = {udata, du, z};
        const double *Xp[3] = {dU, dD, dz};
        udata->ops->nvlinearcombination(3, c, X, w);
        double *a = fetch(dw, NY, ws);
        N_Vector w2 = N_VClone(udata);
        shud_nv_linearcombination(ws, (int64_t)NY, 3, c, Xp, N_VGetDeviceArrayPointer_ShudB200(w2));
        double *b = fetch(N_VGetDeviceArrayPointer_ShudB200(w2), NY, ws);
        check(same(a, b, NY), "nvlinearcombination == shud_nv_linearcombination");
        free(a); free(b);
        udata->ops->nvdotprodmulti(3, du, X, d3);
        shud_nv_dotprodmulti(ws, (int64_t)NY, 3, dD, Xp, e3);
        check(d3[0] == e3[0] && d3[1] == e3[1] && d3[2] == e3[2], "nvdotprodmulti == shud_nv_dotprodmulti");
        N_VDestroy(w2);
    }
    {
        /* nvcloneempty + nvsetarraypointer: caller-owned host storage becomes the mirror and is pushed to the device */
        N_Vector e = udata->ops->nvcloneempty(udata);
        check(e != NULL && N_VGetDeviceArrayPointer_ShudB200(e) == NULL, "nvcloneempty");
        if (e) N_VDestroy(e);
    }
    check(N_VOpsCalled_ShudB200() == 0, "no operation outside the CVODE + SPGMR set was needed");

    /* ---- CVode: the ops-table path and the device-fused path carry the state to the same place ---- */
    double yend[2][4];
    long nst[2], nfe[2], nli[2];
    for (int arm = 0; arm < 2; arm++) {
        mirror = N_VGetArrayPointer(udata);               /* refreshes the mirror first: write after the call */
        for (size_t i = 0; i < NY; i++) mirror[i] = y0[i];
        N_VCopyToDevice_ShudB200(udata);
        shud_b200_prime(gpu, y0);
        shud_cv *cv = NULL;
        if (shud_cv_create(shud_b200_f, gpu, 0.0, udata, &cv) != SHUD_CV_SUCCESS) return 1;
        shud_cv_sstolerances(cv, 1e-4, 1e-4);            /* cvode_config.cpp:162-193 */
        shud_cv_set_init_step(cv, 0.1); shud_cv_set_min_step(cv, 1e-6); shud_cv_set_max_step(cv, 10.0);
        shud_cv_set_max_num_steps(cv, 1000000); shud_cv_set_maxl(cv, 0);
        shud_cv_fused fused = {0};
        if (arm == 1) {
            if (shud_b200_cv_fused_create(gpu, ws, 5, &fused) != SHUD_OK) return 1;
            shud_cv_set_fused(cv, &fused);
        }
        double t = 0.0;
        int flag = 0;
        for (double tout = 10.0; tout <= t_end + 1e-9 && flag >= 0; tout += 10.0) flag = shud_cv_solve(cv, tout, udata, &t, SHUD_CV_NORMAL);
        check(flag >= 0 && fabs(t - floor(t_end / 10.0) * 10.0) < 1e-9, arm ? "shud_cv_solve (device-fused Newton-Krylov)" : "shud_cv_solve (ops table)");
        shud_cv_stats st;
        shud_cv_get_stats(cv, &st);
        nst[arm] = st.nst; nfe[arm] = st.nfe + st.nfeLS; nli[arm] = st.nli;
        const double *s = N_VSummary_ShudB200(udata);              /* Model_Data::summary */
        double a = 0, b = 0, c = 0, d = 0;
        for (size_t i = 0; i < Ne; i++) { a += s[i]; b += s[Ne + i]; c += s[2 * Ne + i]; }
        for (size_t i = 3 * Ne; i < NY; i++) d += s[i];
        yend[arm][0] = a; yend[arm][1] = b; yend[arm][2] = c; yend[arm][3] = d;
        printf("cvode arm %d: t=%.3f nst=%ld rhs_calls=%ld nli=%ld q=%d  sum(Ysurf)=%.12g sum(Yunsat)=%.12g sum(Ygw)=%.12g sum(Yriv+lake)=%.12g\n",
               arm, t, st.nst, st.nfe + st.nfeLS, st.nli, st.qlast, a, b, c, d);
        shud_cv_free(cv);
        if (arm == 1) shud_b200_cv_fused_destroy(&fused);
    }
    {
        int close = 1;
        for (int k = 0; k < 4; k++) close = close && fabs(yend[0][k] - yend[1][k]) <= 1e-6 * (fabs(yend[0][k]) + 1e-3);
        check(close, "both paths end in the same state (1e-6 of the block sums)");
        check(nst[0] > 0 && labs(nst[0] - nst[1]) <= 1 + nst[0] / 10 && nli[1] > 0 && nfe[1] > 0, "both paths take the same steps (within 10 %)");
    }
    N_VDestroy(w); N_VDestroy(z); N_VDestroy(du); N_VDestroy(udata);
    shud_nv_ws_destroy(ws);
    shud_b200_destroy(gpu);
    shud_b200_mesh_free(block);
    printf("%s\n", fails ? "FAILED" : "ALL OK");
    return fails ? 4 : 0;
}
