// shud_nvector_generic.cpp - N_VClone / N_VLinearSum / ... : the generic dispatch through v->ops that
// sundials_nvector.c provides, for host code built without SUNDIALS (include/shud_sundials.h).  Plain host code: it is
// compiled into the CUDA library and into the CPU-only checker library alike.
#include "shud_sundials.h"

extern "C" {
#ifndef SHUD_HAVE_SUNDIALS
// generic dispatch through v->ops, as sundials_nvector.c does it (fused operations fall back to loops over the
// standard ones when a vector does not provide them)
N_Vector N_VClone(N_Vector w) { return w->ops->nvclone(w); }
void N_VDestroy(N_Vector v) { if (v) { if (v->ops && v->ops->nvdestroy) v->ops->nvdestroy(v); } }
realtype *N_VGetArrayPointer(N_Vector v) { return v->ops->nvgetarraypointer(v); }
sunindextype N_VGetLength(N_Vector v) { return v->ops->nvgetlength(v); }
void N_VLinearSum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z) { z->ops->nvlinearsum(a, x, b, y, z); }
void N_VConst(realtype c, N_Vector z) { z->ops->nvconst(c, z); }
void N_VProd(N_Vector x, N_Vector y, N_Vector z) { z->ops->nvprod(x, y, z); }
void N_VDiv(N_Vector x, N_Vector y, N_Vector z) { z->ops->nvdiv(x, y, z); }
void N_VScale(realtype c, N_Vector x, N_Vector z) { z->ops->nvscale(c, x, z); }
void N_VAbs(N_Vector x, N_Vector z) { z->ops->nvabs(x, z); }
void N_VInv(N_Vector x, N_Vector z) { z->ops->nvinv(x, z); }
void N_VAddConst(N_Vector x, realtype b, N_Vector z) { z->ops->nvaddconst(x, b, z); }
realtype N_VDotProd(N_Vector x, N_Vector y) { return y->ops->nvdotprod(x, y); }
realtype N_VMaxNorm(N_Vector x) { return x->ops->nvmaxnorm(x); }
realtype N_VWrmsNorm(N_Vector x, N_Vector w) { return x->ops->nvwrmsnorm(x, w); }
realtype N_VMin(N_Vector x) { return x->ops->nvmin(x); }
int N_VLinearCombination(int nvec, realtype *c, N_Vector *X, N_Vector z) {
    if (z->ops->nvlinearcombination) return z->ops->nvlinearcombination(nvec, c, X, z);
    z->ops->nvscale(c[0], X[0], z);
    for (int i = 1; i < nvec; i++) z->ops->nvlinearsum(c[i], X[i], 1.0, z, z);
    return 0;
}
int N_VScaleAddMulti(int nvec, realtype *a, N_Vector x, N_Vector *Y, N_Vector *Z) {
    if (x->ops->nvscaleaddmulti) return x->ops->nvscaleaddmulti(nvec, a, x, Y, Z);
    for (int i = 0; i < nvec; i++) x->ops->nvlinearsum(a[i], x, 1.0, Y[i], Z[i]);
    return 0;
}
int N_VDotProdMulti(int nvec, N_Vector x, N_Vector *Y, realtype *d) {
    if (x->ops->nvdotprodmulti) return x->ops->nvdotprodmulti(nvec, x, Y, d);
    for (int i = 0; i < nvec; i++) d[i] = x->ops->nvdotprod(x, Y[i]);
    return 0;
}
#endif

}  // extern "C"
