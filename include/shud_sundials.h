/* shud_sundials.h - the SUNDIALS-facing side of the boundary: the device N_Vector as a SUNDIALS 6 N_Vector
 * (SURVEY.md section 8(b), rows "Vector construction", "Host access to vector data", "N_Vector ops table").
 *
 * What the reference binds today and what replaces it:
 *   N_VNew_Serial(NY, sunctx) / N_VNew_OpenMP(NY, nthreads, sunctx)   src/Model/shud.cpp:59-64
 *       -> N_VNew_ShudB200(NY, ws, gpu, sunctx)
 *   CVODE and SUNLinSol_SPGMR clone their work vectors from udata      src/Equations/cvode_config.cpp:169,176
 *       -> v->ops->nvclone / nvdestroy of the table below
 *   NV_Ith_S / NV_DATA_S host access (SetIC2Y, summary, wbdiag)        src/ModelData/MD_initialize.cpp:117-135,
 *                                                                      src/ModelData/MD_update.cpp:190-216, shud.cpp:147
 *       -> N_VGetArrayPointer (host mirror, reference order; refreshed from the device on every call) and
 *          N_VCopyToDevice_ShudB200 after host writes
 *   int f(realtype t, N_Vector y, N_Vector ydot, void *MD)             src/Model/f.hpp:12
 *       -> shud_b200_f (same CVRhsFn shape; user_data = the shud_ctx)
 *
 * SUNDIALS is third-party and not vendored by the reference (configure:17-21 pins cvode-6.0.0).  When the real
 * headers are available compile with -DSHUD_HAVE_SUNDIALS and they are used; otherwise the declarations below
 * restate the SUNDIALS 6.0 generic N_Vector ABI (struct _generic_N_Vector, struct _generic_N_Vector_Ops, member order
 * of sundials/sundials_nvector.h v6.0.0) so that the table can be built, called and tested without SUNDIALS.
 */
#ifndef SHUD_SUNDIALS_H
#define SHUD_SUNDIALS_H
#include <stdint.h>
#include <stdio.h>
#include "shud_b200.h"
#include "shud_nvector.h"

#ifdef SHUD_HAVE_SUNDIALS
#include <sundials/sundials_nvector.h>
#else
#ifdef __cplusplus
extern "C" {
#endif
typedef double realtype;
typedef int64_t sunindextype;
typedef int booleantype;
#define SUNTRUE 1
#define SUNFALSE 0
typedef void *SUNContext; /* opaque here */
typedef enum {
    SUNDIALS_NVEC_SERIAL, SUNDIALS_NVEC_PARALLEL, SUNDIALS_NVEC_OPENMP, SUNDIALS_NVEC_PTHREADS, SUNDIALS_NVEC_PARHYP,
    SUNDIALS_NVEC_PETSC, SUNDIALS_NVEC_CUDA, SUNDIALS_NVEC_HIP, SUNDIALS_NVEC_SYCL, SUNDIALS_NVEC_RAJA,
    SUNDIALS_NVEC_OPENMPDEV, SUNDIALS_NVEC_TRILINOS, SUNDIALS_NVEC_MANYVECTOR, SUNDIALS_NVEC_MPIMANYVECTOR,
    SUNDIALS_NVEC_MPIPLUSX, SUNDIALS_NVEC_CUSTOM
} N_Vector_ID;
typedef struct _generic_N_Vector_Ops *N_Vector_Ops;
typedef struct _generic_N_Vector *N_Vector;
typedef N_Vector *N_Vector_S;
struct _generic_N_Vector_Ops {
    /* constructors, destructors, utility operations */
    N_Vector_ID (*nvgetvectorid)(N_Vector);
    N_Vector (*nvclone)(N_Vector);
    N_Vector (*nvcloneempty)(N_Vector);
    void (*nvdestroy)(N_Vector);
    void (*nvspace)(N_Vector, sunindextype *, sunindextype *);
    realtype *(*nvgetarraypointer)(N_Vector);
    realtype *(*nvgetdevicearraypointer)(N_Vector);
    void (*nvsetarraypointer)(realtype *, N_Vector);
    void *(*nvgetcommunicator)(N_Vector);
    sunindextype (*nvgetlength)(N_Vector);
    /* standard vector operations */
    void (*nvlinearsum)(realtype, N_Vector, realtype, N_Vector, N_Vector);
    void (*nvconst)(realtype, N_Vector);
    void (*nvprod)(N_Vector, N_Vector, N_Vector);
    void (*nvdiv)(N_Vector, N_Vector, N_Vector);
    void (*nvscale)(realtype, N_Vector, N_Vector);
    void (*nvabs)(N_Vector, N_Vector);
    void (*nvinv)(N_Vector, N_Vector);
    void (*nvaddconst)(N_Vector, realtype, N_Vector);
    realtype (*nvdotprod)(N_Vector, N_Vector);
    realtype (*nvmaxnorm)(N_Vector);
    realtype (*nvwrmsnorm)(N_Vector, N_Vector);
    realtype (*nvwrmsnormmask)(N_Vector, N_Vector, N_Vector);
    realtype (*nvmin)(N_Vector);
    realtype (*nvwl2norm)(N_Vector, N_Vector);
    realtype (*nvl1norm)(N_Vector);
    void (*nvcompare)(realtype, N_Vector, N_Vector);
    booleantype (*nvinvtest)(N_Vector, N_Vector);
    booleantype (*nvconstrmask)(N_Vector, N_Vector, N_Vector);
    realtype (*nvminquotient)(N_Vector, N_Vector);
    /* fused vector operations */
    int (*nvlinearcombination)(int, realtype *, N_Vector *, N_Vector);
    int (*nvscaleaddmulti)(int, realtype *, N_Vector, N_Vector *, N_Vector *);
    int (*nvdotprodmulti)(int, N_Vector, N_Vector *, realtype *);
    /* vector array operations */
    int (*nvlinearsumvectorarray)(int, realtype, N_Vector *, realtype, N_Vector *, N_Vector *);
    int (*nvscalevectorarray)(int, realtype *, N_Vector *, N_Vector *);
    int (*nvconstvectorarray)(int, realtype, N_Vector *);
    int (*nvwrmsnormvectorarray)(int, N_Vector *, N_Vector *, realtype *);
    int (*nvwrmsnormmaskvectorarray)(int, N_Vector *, N_Vector *, N_Vector, realtype *);
    int (*nvscaleaddmultivectorarray)(int, int, realtype *, N_Vector *, N_Vector **, N_Vector **);
    int (*nvlinearcombinationvectorarray)(int, int, realtype *, N_Vector **, N_Vector *);
    /* local reduction operations */
    realtype (*nvdotprodlocal)(N_Vector, N_Vector);
    realtype (*nvmaxnormlocal)(N_Vector);
    realtype (*nvminlocal)(N_Vector);
    realtype (*nvl1normlocal)(N_Vector);
    booleantype (*nvinvtestlocal)(N_Vector, N_Vector);
    booleantype (*nvconstrmasklocal)(N_Vector, N_Vector, N_Vector);
    realtype (*nvminquotientlocal)(N_Vector, N_Vector);
    realtype (*nvwsqrsumlocal)(N_Vector, N_Vector);
    realtype (*nvwsqrsummasklocal)(N_Vector, N_Vector, N_Vector);
    /* single buffer reduction operations */
    int (*nvdotprodmultilocal)(int, N_Vector, N_Vector *, realtype *);
    int (*nvdotprodmultiallreduce)(int, N_Vector, realtype *);
    /* XBraid interface operations */
    int (*nvbufsize)(N_Vector, sunindextype *);
    int (*nvbufpack)(N_Vector, void *);
    int (*nvbufunpack)(N_Vector, void *);
    /* debugging functions */
    void (*nvprint)(N_Vector);
    void (*nvprintfile)(N_Vector, FILE *);
};
struct _generic_N_Vector {
    void *content;
    N_Vector_Ops ops;
    SUNContext sunctx;
};
#ifdef __cplusplus
}
#endif
#endif /* SHUD_HAVE_SUNDIALS */

#ifdef __cplusplus
extern "C" {
#endif

/* Allreduce hook of a vector distributed over several GPUs (one partition per rank): called with `n` host doubles
 * that are reduced in place over all ranks; op: 0 sum, 1 max, 2 min.  shud_b200_allreduce (below) is the one the
 * library provides over its own NCCL communicator; NULL = single-GPU vector. */
typedef int (*shud_nv_allreduce_fn)(void *comm, double *vals, int n, int op);

/* content of a SHUD B200 N_Vector (v->content) */
typedef struct shud_nv_content {
    sunindextype length;        /* local length (NY of this partition) */
    sunindextype global_length; /* length of the whole vector (= length on one GPU) */
    int own_dev, own_host;
    double *dev;                /* device data, DEVICE ORDER when gpu != NULL */
    double *host;               /* pinned host mirror, the reference's blocked order (allocated on first use) */
    shud_nvws *ws;              /* reduction workspace + stream the operations run on */
    struct shud_ctx *gpu;       /* RHS context: provides the device-order <-> reference-order permutation (may be NULL) */
    shud_nv_allreduce_fn allreduce;
    void *comm;
} shud_nv_content;

/* N_VNew_Serial / N_VNew_OpenMP replacement (src/Model/shud.cpp:59-64).  `gpu` may be NULL (plain device vector,
 * host mirror in the same order); with a context, length must equal shud_b200_ny(gpu).  NULL on failure. */
N_Vector N_VNew_ShudB200(sunindextype length, shud_nvws *ws, struct shud_ctx *gpu, SUNContext sunctx);
/* wrap existing device storage (not owned) */
N_Vector N_VMake_ShudB200(sunindextype length, double *dev, shud_nvws *ws, struct shud_ctx *gpu, SUNContext sunctx);
/* distributed vector: global length + allreduce hook (inherited by clones).  global_length = the number of OWNED
 * entries over all ranks: a partition whose rivers are cut carries ghost entries in its local vector (ydot = 0, state
 * from the exchange); they are zero in every correction, residual and Krylov vector, add nothing to a sum and must not
 * add to N.  fn == shud_b200_nv_allreduce with comm = the shud_ctx: the library reduces on the device - inside the
 * reduction kernels over NVLink mailboxes after shud_b200_p2p_connect, else ncclAllReduce on the stream. */
void N_VSetDistributed_ShudB200(N_Vector v, sunindextype global_length, shud_nv_allreduce_fn fn, void *comm);
/* host mirror -> device (after SetIC2Y-style writes through N_VGetArrayPointer); device -> host mirror (what
 * N_VGetArrayPointer does implicitly).  With a context the mirror is in the reference's order, the device vector in
 * device order.  N_VSummary_ShudB200: Model_Data::summary semantics (BC heads / stages replace the frozen rows). */
int N_VCopyToDevice_ShudB200(N_Vector v);
int N_VCopyFromDevice_ShudB200(N_Vector v);
double *N_VSummary_ShudB200(N_Vector v);
double *N_VGetDeviceArrayPointer_ShudB200(N_Vector v);
/* how many times an operation of the table the CVODE+SPGMR configuration never uses was called (they are
 * implemented, the counter exists for the integration test) */
long N_VOpsCalled_ShudB200(void);

/* the allreduce hook over the NCCL communicator of an RHS context (shud_b200_comm_init): comm = the shud_ctx */
int shud_b200_nv_allreduce(void *comm, double *vals, int n, int op);

/* the reference's f() (src/Model/f.hpp:12, f.cpp:2-32): CVRhsFn on two SHUD B200 vectors; user_data = shud_ctx*.
 * Returns 0, or -1 (unrecoverable, what CVODE expects) when the device error word is set - checked only when
 * shud_b200_f_check(1) was requested, since the check costs a stream synchronisation. */
int shud_b200_f(realtype t, N_Vector y, N_Vector ydot, void *user_data);
int shud_b200_f_exchange(realtype t, N_Vector y, N_Vector ydot, void *user_data); /* partition: halo exchange + RHS */
void shud_b200_f_check(int on);

/* generic dispatch (what sundials_nvector.c provides), for host code built without SUNDIALS */
#ifndef SHUD_HAVE_SUNDIALS
N_Vector N_VClone(N_Vector w);
void N_VDestroy(N_Vector v);
realtype *N_VGetArrayPointer(N_Vector v);
sunindextype N_VGetLength(N_Vector v);
void N_VLinearSum(realtype a, N_Vector x, realtype b, N_Vector y, N_Vector z);
void N_VConst(realtype c, N_Vector z);
void N_VProd(N_Vector x, N_Vector y, N_Vector z);
void N_VDiv(N_Vector x, N_Vector y, N_Vector z);
void N_VScale(realtype c, N_Vector x, N_Vector z);
void N_VAbs(N_Vector x, N_Vector z);
void N_VInv(N_Vector x, N_Vector z);
void N_VAddConst(N_Vector x, realtype b, N_Vector z);
realtype N_VDotProd(N_Vector x, N_Vector y);
realtype N_VMaxNorm(N_Vector x);
realtype N_VWrmsNorm(N_Vector x, N_Vector w);
realtype N_VMin(N_Vector x);
int N_VLinearCombination(int nvec, realtype *c, N_Vector *X, N_Vector z);
int N_VScaleAddMulti(int nvec, realtype *a, N_Vector x, N_Vector *Y, N_Vector *Z);
int N_VDotProdMulti(int nvec, N_Vector x, N_Vector *Y, realtype *dotprods);
#endif

#ifdef __cplusplus
}
#endif
#endif
