// shud_nvec.cu - device N_Vector arithmetic for CVODE(BDF, Newton) + CVLS + SPGMR on the SHUD state
// vectors (include/shud_nvector.h).  Pure HBM streaming: grid-stride kernels with 4 independent 8-byte
// loads per thread and array in flight, grid sized as a multiple of the 148 SMs.  Reductions: per-thread
// partial -> fixed-shape warp-shuffle tree -> one partial per block -> the last block to finish sums the
// block partials in index order (no floating-point atomics: the result is run-to-run reproducible) and
// stores the scalar(s) straight into mapped pinned host memory, so the host needs one stream
// synchronisation and no extra copy.
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <atomic>
#include <cstring>
#include <cstdlib>

#include "shud_b200.h"
#include "shud_nvector.h"

namespace {

constexpr int NT = 256;          // threads per block
constexpr int MAXB = 148 * 8;    // blocks: 8 per SM
constexpr int UNROLL = 4;

struct Ptrs {
    const double *p[SHUD_NV_MAXVEC];
};
struct MPtrs {
    double *p[SHUD_NV_MAXVEC];
};
struct Coef {
    double c[SHUD_NV_MAXVEC];
};

inline int grid_for(int64_t n, int cap = MAXB) {
    int64_t b = (n + (int64_t)NT * UNROLL - 1) / ((int64_t)NT * UNROLL);
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}
// blocks of a reduction: 4 per SM (half as many partial sums, atomics and fences in the kernel's tail as with 8; measured
// on the Newton-Krylov step: 8 -> 0.755, 4 -> 0.736, 2 -> 0.829 ms); A/B knobs SHUD_NV_RBLOCKS / SHUD_NV_MBLOCKS
inline int rgrid_for(int64_t n) {
    static const int per_sm = [] { const char *e = getenv("SHUD_NV_RBLOCKS"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    return grid_for(n, 148 * per_sm);
}
inline int mgrid_for(int64_t n) {
    static const int per_sm = [] { const char *e = getenv("SHUD_NV_MBLOCKS"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    return grid_for(n, 148 * per_sm);
}

// ---------------- streaming ----------------
template <class F>
__global__ void __launch_bounds__(NT) k_map(int64_t n, F f) {
    const int64_t stride = (int64_t)gridDim.x * NT;
    int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) f(i + u * stride);
    }
    for (; i < n; i += stride) f(i);
}

struct FLinearSum {
    double a, b; const double *x, *y; double *z;
    __device__ void operator()(int64_t i) const { z[i] = a * x[i] + b * y[i]; }
};
struct FAxpy {  // y += a x  (in-place form CVODE uses most: one stream fewer)
    double a; const double *x; double *y;
    __device__ void operator()(int64_t i) const { y[i] = a * x[i] + y[i]; }
};
struct FConst { double c; double *z; __device__ void operator()(int64_t i) const { z[i] = c; } };
struct FProd { const double *x, *y; double *z; __device__ void operator()(int64_t i) const { z[i] = x[i] * y[i]; } };
struct FDiv { const double *x, *y; double *z; __device__ void operator()(int64_t i) const { z[i] = x[i] / y[i]; } };
struct FScale { double c; const double *x; double *z; __device__ void operator()(int64_t i) const { z[i] = c * x[i]; } };
struct FAbs { const double *x; double *z; __device__ void operator()(int64_t i) const { z[i] = fabs(x[i]); } };
struct FInv { const double *x; double *z; __device__ void operator()(int64_t i) const { z[i] = 1.0 / x[i]; } };
struct FAddConst { const double *x; double b; double *z; __device__ void operator()(int64_t i) const { z[i] = x[i] + b; } };
struct FCompare { double c; const double *x; double *z; __device__ void operator()(int64_t i) const { z[i] = (fabs(x[i]) >= c) ? 1.0 : 0.0; } };
// The linear combinations are instantiated per vector count: with the count known the loads of all addends are
// issued before the first multiply-add (with a run-time loop each load waited for the sum of the ones before it and the
// kernels ran at 2.9 TB/s); the sum itself keeps its left-to-right order.
template <int NV>
__device__ __forceinline__ double lincomb_at(const Coef &c, const Ptrs &X, int64_t i) {
    double x[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = X.p[k][i];
    double s = c.c[0] * x[0];
#pragma unroll
    for (int k = 1; k < NV; k++) s += c.c[k] * x[k];
    return s;
}
template <int NV>
struct FLinComb {
    Coef c; Ptrs X; double *z;
    __device__ void operator()(int64_t i) const { z[i] = lincomb_at<NV>(c, X, i); }
};
struct FScaleAddMulti {
    int nv; Coef a; const double *x; Ptrs Y; MPtrs Z;
    __device__ void operator()(int64_t i) const {
        const double xi = x[i];
        for (int k = 0; k < nv; k++) Z.p[k][i] = a.c[k] * xi + Y.p[k][i];
    }
};
struct FLinSumVA {
    int nv; double a, b; Ptrs X, Y; MPtrs Z;
    __device__ void operator()(int64_t i) const { for (int k = 0; k < nv; k++) Z.p[k][i] = a * X.p[k][i] + b * Y.p[k][i]; }
};
struct FScaleVA {
    int nv; Coef c; Ptrs X; MPtrs Z;
    __device__ void operator()(int64_t i) const { for (int k = 0; k < nv; k++) Z.p[k][i] = c.c[k] * X.p[k][i]; }
};
struct FDqPerturb {
    double sigma; const double *vs, *ewt, *y; double *yt;
    __device__ void operator()(int64_t i) const { yt[i] = sigma * (vs[i] / ewt[i]) + y[i]; }
};
struct FDqCombine {
    double sigma, gamma; const double *vs, *ewt, *fp, *fy; double *out;
    __device__ void operator()(int64_t i) const {
        const double w = ewt[i], v = vs[i] / w;
        const double jv = (1.0 / sigma) * fp[i] + (-1.0 / sigma) * fy[i];
        out[i] = w * (v + (-gamma) * jv);
    }
};
struct FEwt {
    double rtol, atol; const double *y; double *w;
    __device__ void operator()(int64_t i) const { w[i] = 1.0 / (rtol * fabs(y[i]) + atol); }
};
struct TEwtNorm {  // w = 1 / (rtol |y| + atol) (FEwt); term = (y w)^2
    double rtol, atol; const double *y; double *w;
    __device__ double term(int, int64_t i) const {
        const double yi = y[i], wi = 1.0 / (rtol * fabs(yi) + atol);
        w[i] = wi;
        const double t = yi * wi;
        return t * t;
    }
};
struct FNewtonResid {
    double gamma; const double *f, *psi, *y; double *r;
    __device__ void operator()(int64_t i) const { r[i] = (gamma * f[i] + psi[i]) - y[i]; }
};
struct FConstVA {
    int nv; double c; MPtrs Z;
    __device__ void operator()(int64_t i) const { for (int k = 0; k < nv; k++) Z.p[k][i] = c; }
};

// ---------------- reductions ----------------
enum { R_SUM = 0, R_MAX = 1, R_MIN = 2 };
template <int KIND>
__device__ __forceinline__ double comb(double a, double b) {
    if (KIND == R_SUM) return a + b;
    if (KIND == R_MAX) return a < b ? b : a;
    return a > b ? b : a;
}
template <int KIND>
__device__ __forceinline__ double ident() { return KIND == R_SUM ? 0.0 : (KIND == R_MAX ? 0.0 : DBL_MAX); }

template <int KIND>
__device__ __forceinline__ double block_reduce(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = comb<KIND>(v, __shfl_down_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = ident<KIND>();
    if (threadIdx.x < 32) {
        t = (threadIdx.x < NT / 32) ? sm[threadIdx.x] : ident<KIND>();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = comb<KIND>(t, __shfl_down_sync(0xffffffffu, t, o));
    }
    __syncthreads();
    return t;  // valid in thread 0
}

// ---- allreduce over NVLink inside the reduction kernel (one process per GPU, the ranks' mailboxes mapped through
// CUDA IPC by shud_b200_p2p_connect): the last block of the reduction stores the rank's partial results into every
// rank's mailbox (peer stores + a release of the sequence number), waits for the other ranks' sequence numbers in its
// own mailbox, and combines the partials in rank order - the same bits on every rank, no second kernel, no NCCL call.
// Two slots by the parity of the sequence number: a rank can be at most one reduction ahead of the slowest one.
struct ArBox {
    double val[2][SHUD_NV_MAXRANKS][SHUD_NV_MAXVEC];
    unsigned long long tag[2][SHUD_NV_MAXRANKS];
};
static_assert(sizeof(ArBox) <= SHUD_NV_ARBOX_BYTES, "mailbox larger than the region shud_b200_p2p_export reserves");
struct PeerAR {
    int nranks, rank;
    unsigned long long seq;
    ArBox *box[SHUD_NV_MAXRANKS];
};
__device__ __forceinline__ void ar_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ar_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// NV accumulators per thread (1 for plain reductions, up to SHUD_NV_MAXVEC for the multi forms);
// F::term(k, i) is the value element i contributes to accumulator k.
// post: 0 none, 1 sqrt(v / nglob), 2 sqrt(v)
template <int KIND, int NV, class F>
__global__ void __launch_bounds__(NT) k_reduce(int64_t n, F f, int nv, double *partial, unsigned *counter,
                                               double *d_out, volatile double *h_out, int post, double nglob,
                                               volatile double *h_ticket = nullptr, double ticket = 0.0,
                                               PeerAR P = PeerAR{}) {
    __shared__ double sv[NV];
    __shared__ double sm[NT / 32];
    __shared__ bool last;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = ident<KIND>();
    const int64_t stride = (int64_t)gridDim.x * NT;
    int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int k = 0; k < NV; k++)
                if (k < nv) acc[k] = comb<KIND>(acc[k], f.term(k, i + u * stride));
    }
    for (; i < n; i += stride)
#pragma unroll
        for (int k = 0; k < NV; k++)
            if (k < nv) acc[k] = comb<KIND>(acc[k], f.term(k, i));
#pragma unroll
    for (int k = 0; k < NV; k++) {
        if (k < nv) {
            const double b = block_reduce<KIND>(acc[k], sm);
            if (threadIdx.x == 0) partial[(size_t)k * MAXB + blockIdx.x] = b;
        }
    }
    __threadfence();
    if (threadIdx.x == 0) last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int k = 0; k < nv; k++) {
        double v = ident<KIND>();
        for (int b = threadIdx.x; b < (int)gridDim.x; b += NT) v = comb<KIND>(v, partial[(size_t)k * MAXB + b]);
        v = block_reduce<KIND>(v, sm);
        if (threadIdx.x == 0) sv[k] = v;
    }
    if (P.nranks > 1) {
        __syncthreads();
        const int par = (int)(P.seq & 1ull);
        ArBox *const me = P.box[P.rank];
        if ((int)threadIdx.x < P.nranks) {
            ArBox *const dst = P.box[threadIdx.x];  // thread r serves rank r: send to it, then wait for it
            for (int k = 0; k < nv; k++) dst->val[par][P.rank][k] = sv[k];
            __threadfence_system();
            ar_release(&dst->tag[par][P.rank], P.seq);
            const long long t0 = clock64();
            while (ar_acquire(&me->tag[par][threadIdx.x]) < P.seq) {
                if (clock64() - t0 > 6000000000ll) {  // ~3 s: a rank is missing - poison the result instead of hanging
                    for (int k = 0; k < nv; k++) ((volatile double *)me->val[par][threadIdx.x])[k] = nan("");
                    break;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 0; k < nv; k++) {
                double v = ident<KIND>();
                for (int r = 0; r < P.nranks; r++) v = comb<KIND>(v, ((const volatile double *)me->val[par][r])[k]);
                sv[k] = v;
            }
        }
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < nv; k++) {
            double v = sv[k];
            if (post == 1) v = sqrt(v / nglob);
            else if (post == 2) v = sqrt(v);
            d_out[k] = v;
            h_out[k] = v;
        }
    }
    // the host spins on this word instead of synchronising the stream: the results above (mapped host memory) are
    // ordered ahead of it
    if (h_ticket && threadIdx.x == 0) { __threadfence_system(); *h_ticket = ticket; }
}

struct TDot { const double *x, *y; __device__ double term(int, int64_t i) const { return x[i] * y[i]; } };
struct TAbs { const double *x; __device__ double term(int, int64_t i) const { return fabs(x[i]); } };
struct TVal { const double *x; __device__ double term(int, int64_t i) const { return x[i]; } };
struct TWSqr { const double *x, *w; __device__ double term(int, int64_t i) const { const double t = x[i] * w[i]; return t * t; } };
struct TWSqrMask {
    const double *x, *w, *id;
    __device__ double term(int, int64_t i) const { const double t = x[i] * w[i]; return id[i] > 0.0 ? t * t : 0.0; }
};
struct TMinQuot {
    const double *num, *den;
    __device__ double term(int, int64_t i) const { return den[i] != 0.0 ? num[i] / den[i] : DBL_MAX; }
};
struct TNewtonUpdate {  // y += x, acor += x as a side effect; term = (x w)^2
    const double *x, *w; double *y, *acor;
    __device__ double term(int, int64_t i) const {
        const double xi = x[i];
        y[i] = y[i] + xi;
        acor[i] = acor[i] + xi;
        const double t = xi * w[i];
        return t * t;
    }
};
struct TDotMulti { const double *x; Ptrs Y; __device__ double term(int k, int64_t i) const { return x[i] * Y.p[k][i]; } };
struct TWSqrMulti {
    Ptrs X, W;
    __device__ double term(int k, int64_t i) const { const double t = X.p[k][i] * W.p[k][i]; return t * t; }
};
// map + flag reductions (z written as a side effect; term = 1 where the test fails)
struct TInvTest {
    const double *x; double *z;
    __device__ double term(int, int64_t i) const {
        const double v = x[i];
        if (v == 0.0) return 1.0;
        z[i] = 1.0 / v;
        return 0.0;
    }
};
struct TConstrMask {
    const double *c, *x; double *m;
    __device__ double term(int, int64_t i) const {
        const double ci = c[i], xi = x[i];
        m[i] = 0.0;
        if (ci == 0.0) return 0.0;
        // |c| = 2: strict sign; |c| = 1: non-strict (SUNDIALS N_VConstrMask)
        const bool bad = (fabs(ci) > 1.5) ? (xi * ci <= 0.0) : (xi * ci < 0.0);
        if (bad) { m[i] = 1.0; return 1.0; }
        return 0.0;
    }
};

}  // namespace

struct shud_nvws {
    int device;
    cudaStream_t stream;
    double *partial;       // [SHUD_NV_MAXVEC][MAXB]
    unsigned *counter;
    double *d_out;         // [SHUD_NV_MAXVEC]
    double *h_out;         // mapped pinned, [SHUD_NV_MAXVEC] results + [1] ticket of the last signalled reduction
    double *h_out_dev;     // device alias of h_out
    double ticket = 0.0;   // last ticket handed out
    // allreduce inside the reduction kernels over the ranks' mailboxes (shud_nv_ws_set_peer_allreduce); preferred to ar_dev
    PeerAR peer{};
    unsigned long long peer_seq = 0;
    // distributed vector: the partial results of a reduction stay on the device, are reduced over the ranks in place
    // (ncclAllReduce on the same stream) and only then copied to the host: one synchronisation per reduction
    shud_nv_allreduce_dev_fn ar_dev = nullptr;
    void *ar_ctx = nullptr;
    int ar_off = 0;        // > 0: reductions are local for the moment (the *local members of the operations table)
};

#define CKN(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            fprintf(stderr, "[shud_nvec] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SHUD_ERR_CUDA;                                                                          \
        }                                                                                                  \
    } while (0)

namespace {
template <class F>
int run_map(shud_nvws *ws, int64_t n, F f) {
    if (!ws) return SHUD_ERR_ARG;
    if (n <= 0) return SHUD_OK;
    k_map<<<mgrid_for(n), NT, 0, ws->stream>>>(n, f);
    CKN(cudaGetLastError());
    return SHUD_OK;
}
// Wait for the reduction that carries `ticket`: spin on the mapped word (a few microseconds sooner than a stream
// synchronisation, and kernels queued behind the reduction keep running); the stream is polled now and then, so an
// error or a lost ticket ends in the ordinary synchronisation.
int wait_ticket(shud_nvws *ws, double ticket) {
    volatile double *p = ws->h_out + SHUD_NV_MAXVEC;
    for (unsigned spin = 1;; spin++) {
        if (*p == ticket) { std::atomic_thread_fence(std::memory_order_acquire); return SHUD_OK; }
        if ((spin & 4095u) == 0 && cudaStreamQuery(ws->stream) != cudaErrorNotReady) break;
    }
    CKN(cudaStreamSynchronize(ws->stream));
    std::atomic_thread_fence(std::memory_order_acquire);
    return *p == ticket ? SHUD_OK : SHUD_ERR_CUDA;
}

template <int KIND, int NV, class F>
int run_reduce(shud_nvws *ws, int64_t n, F f, int nv, int post, double nglob, double *out) {
    if (!ws || !out || nv < 1 || nv > NV) return SHUD_ERR_ARG;
    if (n <= 0) {
        for (int k = 0; k < nv; k++) out[k] = (KIND == R_MIN) ? DBL_MAX : 0.0;
        return SHUD_OK;
    }
    if (ws->peer.nranks > 1 && !ws->ar_off) {
        // every rank calls this with the same nv (SPMD): the kernel itself combines the ranks' partials over NVLink
        PeerAR P = ws->peer;
        P.seq = ++ws->peer_seq;
        const double ticket = (ws->ticket += 1.0);
        k_reduce<KIND, NV, F><<<rgrid_for(n), NT, 0, ws->stream>>>(n, f, nv, ws->partial, ws->counter, ws->d_out, ws->h_out_dev,
                                                                  post, nglob, ws->h_out_dev + SHUD_NV_MAXVEC, ticket, P);
        CKN(cudaGetLastError());
        const int rc = wait_ticket(ws, ticket);
        if (rc) return rc;
        for (int k = 0; k < nv; k++) out[k] = ws->h_out[k];
        return SHUD_OK;
    }
    if (ws->ar_dev && !ws->ar_off) {
        // every rank calls this with the same nv (SPMD): raw partials -> allreduce on the device -> host, post on the host
        k_reduce<KIND, NV, F><<<rgrid_for(n), NT, 0, ws->stream>>>(n, f, nv, ws->partial, ws->counter, ws->d_out,
                                                                  ws->h_out_dev, 0, 1.0);
        CKN(cudaGetLastError());
        if (ws->ar_dev(ws->ar_ctx, ws->d_out, nv, KIND, (void *)ws->stream) != 0) return SHUD_ERR_CUDA;
        CKN(cudaMemcpyAsync(ws->h_out, ws->d_out, sizeof(double) * nv, cudaMemcpyDeviceToHost, ws->stream));
        CKN(cudaStreamSynchronize(ws->stream));
        for (int k = 0; k < nv; k++) {
            const double v = ws->h_out[k];
            out[k] = post == 1 ? sqrt(v / nglob) : (post == 2 ? sqrt(v) : v);
        }
        return SHUD_OK;
    }
    const double ticket = (ws->ticket += 1.0);
    k_reduce<KIND, NV, F><<<rgrid_for(n), NT, 0, ws->stream>>>(n, f, nv, ws->partial, ws->counter, ws->d_out,
                                                              ws->h_out_dev, post, nglob, ws->h_out_dev + SHUD_NV_MAXVEC, ticket);
    CKN(cudaGetLastError());
    const int rc = wait_ticket(ws, ticket);
    if (rc) return rc;
    for (int k = 0; k < nv; k++) out[k] = ws->h_out[k];
    return SHUD_OK;
}
// same reduction, result left in device memory (d_result), no synchronisation: for device-resident
// solver loops (shud_spgmr_solve) that consume the scalar in a following kernel
// h_result (device alias of mapped host memory, optional): the scalar lands in host memory as well; ticket > 0: the
// host may wait_ticket() for it
template <int KIND, class F>
int run_reduce_dev(shud_nvws *ws, int64_t n, F f, double *d_result, double *h_result = nullptr, double ticket = 0.0) {
    if (!ws || !d_result || n <= 0) return SHUD_ERR_ARG;
    PeerAR P{};
    if (ws->peer.nranks > 1 && !ws->ar_off) { P = ws->peer; P.seq = ++ws->peer_seq; }
    k_reduce<KIND, 1, F><<<rgrid_for(n), NT, 0, ws->stream>>>(n, f, 1, ws->partial, ws->counter, d_result,
                                                             h_result ? h_result : ws->h_out_dev + (SHUD_NV_MAXVEC - 1), 0, 1.0,
                                                             ticket > 0.0 ? ws->h_out_dev + SHUD_NV_MAXVEC : nullptr, ticket, P);
    CKN(cudaGetLastError());
    if (P.nranks > 1) return SHUD_OK;  // reduced over the ranks by the kernel itself
    // distributed vector: the scalar is reduced over the ranks where it lies, on the same stream, before the next
    // kernel reads it (every rank runs the same sequence)
    if (ws->ar_dev && !ws->ar_off && ws->ar_dev(ws->ar_ctx, d_result, 1, KIND, (void *)ws->stream) != 0) return SHUD_ERR_CUDA;
    return SHUD_OK;
}
bool fill(Ptrs &P, const double *const *X, int nv) {
    if (!X || nv < 1 || nv > SHUD_NV_MAXVEC) return false;
    for (int k = 0; k < nv; k++) P.p[k] = X[k];
    return true;
}
bool fill(MPtrs &P, double *const *X, int nv) {
    if (!X || nv < 1 || nv > SHUD_NV_MAXVEC) return false;
    for (int k = 0; k < nv; k++) P.p[k] = X[k];
    return true;
}
}  // namespace

extern "C" {

int shud_nv_ws_create(int device, void *stream, shud_nvws **out) {
    if (!out) return SHUD_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) return SHUD_ERR_NO_DEVICE;
    CKN(cudaSetDevice(device));
    shud_nvws *ws = new shud_nvws();
    ws->device = device;
    ws->stream = (cudaStream_t)stream;
    CKN(cudaMalloc(&ws->partial, sizeof(double) * SHUD_NV_MAXVEC * MAXB));
    CKN(cudaMalloc(&ws->counter, sizeof(unsigned)));
    CKN(cudaMemset(ws->counter, 0, sizeof(unsigned)));
    CKN(cudaMalloc(&ws->d_out, sizeof(double) * SHUD_NV_MAXVEC));
    CKN(cudaHostAlloc(&ws->h_out, sizeof(double) * (SHUD_NV_MAXVEC + 1), cudaHostAllocMapped));
    memset(ws->h_out, 0, sizeof(double) * (SHUD_NV_MAXVEC + 1));
    CKN(cudaHostGetDevicePointer((void **)&ws->h_out_dev, ws->h_out, 0));
    *out = ws;
    return SHUD_OK;
}
void shud_nv_ws_destroy(shud_nvws *ws) {
    if (!ws) return;
    cudaSetDevice(ws->device);
    cudaDeviceSynchronize();  // not the stream: its owner may have destroyed it already
    cudaFree(ws->partial); cudaFree(ws->counter); cudaFree(ws->d_out); cudaFreeHost(ws->h_out);
    delete ws;
}

int shud_nv_ws_set_allreduce(shud_nvws *ws, shud_nv_allreduce_dev_fn fn, void *ctx) {
    if (!ws) return SHUD_ERR_ARG;
    ws->ar_dev = fn; ws->ar_ctx = ctx; ws->ar_off = 0;
    return SHUD_OK;
}
int shud_nv_ws_set_peer_allreduce(shud_nvws *ws, int nranks, int rank, void *const *boxes) {
    if (!ws || nranks < 0 || nranks > SHUD_NV_MAXRANKS || rank < 0 || (nranks > 0 && (rank >= nranks || !boxes))) return SHUD_ERR_ARG;
    // the same mailboxes again (every vector of a run is made distributed on the one workspace): the sequence numbers
    // run on - the tags in the mailboxes do
    bool same = nranks > 0 && ws->peer.nranks == nranks && ws->peer.rank == rank;
    for (int r = 0; r < nranks && same; r++) same = ws->peer.box[r] == (ArBox *)boxes[r];
    if (same) return SHUD_OK;
    ws->peer = PeerAR{};
    ws->peer.nranks = nranks; ws->peer.rank = rank;
    for (int r = 0; r < nranks; r++) ws->peer.box[r] = (ArBox *)boxes[r];
    ws->peer_seq = 0;  // fresh mailboxes are zeroed (shud_b200_p2p_export / _connect)
    return SHUD_OK;
}
void shud_nv_ws_local(shud_nvws *ws, int on) { if (ws) ws->ar_off += on ? 1 : -1; }
void *shud_nv_ws_stream(const shud_nvws *ws) { return ws ? (void *)ws->stream : nullptr; }
int shud_nv_ws_device(const shud_nvws *ws) { return ws ? ws->device : -1; }

int shud_nv_linearsum(shud_nvws *ws, int64_t n, double a, const double *x, double b, const double *y, double *z) {
    if (b == 1.0 && z == y) return run_map(ws, n, FAxpy{a, x, z});   // y += a x
    if (a == 1.0 && z == x) return run_map(ws, n, FAxpy{b, y, z});   // x += b y
    return run_map(ws, n, FLinearSum{a, b, x, y, z});
}
int shud_nv_const(shud_nvws *ws, int64_t n, double c, double *z) { return run_map(ws, n, FConst{c, z}); }
int shud_nv_prod(shud_nvws *ws, int64_t n, const double *x, const double *y, double *z) { return run_map(ws, n, FProd{x, y, z}); }
int shud_nv_div(shud_nvws *ws, int64_t n, const double *x, const double *y, double *z) { return run_map(ws, n, FDiv{x, y, z}); }
int shud_nv_scale(shud_nvws *ws, int64_t n, double c, const double *x, double *z) { return run_map(ws, n, FScale{c, x, z}); }
int shud_nv_abs(shud_nvws *ws, int64_t n, const double *x, double *z) { return run_map(ws, n, FAbs{x, z}); }
int shud_nv_inv(shud_nvws *ws, int64_t n, const double *x, double *z) { return run_map(ws, n, FInv{x, z}); }
int shud_nv_addconst(shud_nvws *ws, int64_t n, const double *x, double b, double *z) { return run_map(ws, n, FAddConst{x, b, z}); }
int shud_nv_compare(shud_nvws *ws, int64_t n, double c, const double *x, double *z) { return run_map(ws, n, FCompare{c, x, z}); }

int shud_nv_dotprod(shud_nvws *ws, int64_t n, const double *x, const double *y, double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TDot{x, y}, 1, 0, 1.0, out);
}
int shud_nv_maxnorm(shud_nvws *ws, int64_t n, const double *x, double *out) {
    return run_reduce<R_MAX, 1>(ws, n, TAbs{x}, 1, 0, 1.0, out);
}
int shud_nv_min(shud_nvws *ws, int64_t n, const double *x, double *out) {
    return run_reduce<R_MIN, 1>(ws, n, TVal{x}, 1, 0, 1.0, out);
}
int shud_nv_l1norm(shud_nvws *ws, int64_t n, const double *x, double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TAbs{x}, 1, 0, 1.0, out);
}
int shud_nv_wsqrsum(shud_nvws *ws, int64_t n, const double *x, const double *w, double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TWSqr{x, w}, 1, 0, 1.0, out);
}
int shud_nv_wsqrsum_mask(shud_nvws *ws, int64_t n, const double *x, const double *w, const double *id, double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TWSqrMask{x, w, id}, 1, 0, 1.0, out);
}
int shud_nv_wrmsnorm(shud_nvws *ws, int64_t n, const double *x, const double *w, int64_t ng, double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TWSqr{x, w}, 1, 1, (double)(ng > 0 ? ng : n), out);
}
int shud_nv_wrmsnorm_mask(shud_nvws *ws, int64_t n, const double *x, const double *w, const double *id, int64_t ng,
                          double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TWSqrMask{x, w, id}, 1, 1, (double)(ng > 0 ? ng : n), out);
}
int shud_nv_wl2norm(shud_nvws *ws, int64_t n, const double *x, const double *w, double *out) {
    return run_reduce<R_SUM, 1>(ws, n, TWSqr{x, w}, 1, 2, 1.0, out);
}
int shud_nv_invtest(shud_nvws *ws, int64_t n, const double *x, double *z, int *ok) {
    double bad = 0.;
    int rc = run_reduce<R_SUM, 1>(ws, n, TInvTest{x, z}, 1, 0, 1.0, &bad);
    if (ok) *ok = (bad == 0.0);
    return rc;
}
int shud_nv_constrmask(shud_nvws *ws, int64_t n, const double *c, const double *x, double *m, int *ok) {
    double bad = 0.;
    int rc = run_reduce<R_SUM, 1>(ws, n, TConstrMask{c, x, m}, 1, 0, 1.0, &bad);
    if (ok) *ok = (bad == 0.0);
    return rc;
}
int shud_nv_minquotient(shud_nvws *ws, int64_t n, const double *num, const double *den, double *out) {
    return run_reduce<R_MIN, 1>(ws, n, TMinQuot{num, den}, 1, 0, 1.0, out);
}

int shud_nv_linearcombination(shud_nvws *ws, int64_t n, int nv, const double *c, const double *const *X, double *z) {
    Coef cf; Ptrs P;
    if (!c || !fill(P, X, nv)) return SHUD_ERR_ARG;
    for (int k = 0; k < nv; k++) cf.c[k] = c[k];
    switch (nv) {
        case 1: return run_map(ws, n, FLinComb<1>{cf, P, z});
        case 2: return run_map(ws, n, FLinComb<2>{cf, P, z});
        case 3: return run_map(ws, n, FLinComb<3>{cf, P, z});
        case 4: return run_map(ws, n, FLinComb<4>{cf, P, z});
        case 5: return run_map(ws, n, FLinComb<5>{cf, P, z});
        case 6: return run_map(ws, n, FLinComb<6>{cf, P, z});
        case 7: return run_map(ws, n, FLinComb<7>{cf, P, z});
        default: return run_map(ws, n, FLinComb<8>{cf, P, z});
    }
}
int shud_nv_scaleaddmulti(shud_nvws *ws, int64_t n, int nv, const double *a, const double *x, const double *const *Y,
                          double *const *Z) {
    FScaleAddMulti f; f.nv = nv; f.x = x;
    if (!a || !fill(f.Y, Y, nv) || !fill(f.Z, Z, nv)) return SHUD_ERR_ARG;
    for (int k = 0; k < nv; k++) f.a.c[k] = a[k];
    return run_map(ws, n, f);
}
int shud_nv_dotprodmulti(shud_nvws *ws, int64_t n, int nv, const double *x, const double *const *Y, double *out) {
    TDotMulti f; f.x = x;
    if (!fill(f.Y, Y, nv)) return SHUD_ERR_ARG;
    return run_reduce<R_SUM, SHUD_NV_MAXVEC>(ws, n, f, nv, 0, 1.0, out);
}
int shud_nv_linearsumvectorarray(shud_nvws *ws, int64_t n, int nv, double a, const double *const *X, double b,
                                 const double *const *Y, double *const *Z) {
    FLinSumVA f; f.nv = nv; f.a = a; f.b = b;
    if (!fill(f.X, X, nv) || !fill(f.Y, Y, nv) || !fill(f.Z, Z, nv)) return SHUD_ERR_ARG;
    return run_map(ws, n, f);
}
int shud_nv_scalevectorarray(shud_nvws *ws, int64_t n, int nv, const double *c, const double *const *X,
                             double *const *Z) {
    FScaleVA f; f.nv = nv;
    if (!c || !fill(f.X, X, nv) || !fill(f.Z, Z, nv)) return SHUD_ERR_ARG;
    for (int k = 0; k < nv; k++) f.c.c[k] = c[k];
    return run_map(ws, n, f);
}
int shud_nv_constvectorarray(shud_nvws *ws, int64_t n, int nv, double c, double *const *Z) {
    FConstVA f; f.nv = nv; f.c = c;
    if (!fill(f.Z, Z, nv)) return SHUD_ERR_ARG;
    return run_map(ws, n, f);
}
int shud_nv_ewt(shud_nvws *ws, int64_t n, double rtol, double atol, const double *y, double *ewt) {
    return run_map(ws, n, FEwt{rtol, atol, y, ewt});
}
int shud_nv_ewt_wrms(shud_nvws *ws, int64_t n, double rtol, double atol, const double *y, double *ewt, int64_t ng, double *nrm) {
    return run_reduce<R_SUM, 1>(ws, n, TEwtNorm{rtol, atol, y, ewt}, 1, 1, (double)(ng > 0 ? ng : n), nrm);
}
int shud_nv_newton_resid(shud_nvws *ws, int64_t n, double gamma, const double *f, const double *psi, const double *y,
                         double *r) {
    return run_map(ws, n, FNewtonResid{gamma, f, psi, y, r});
}
int shud_nv_newton_update(shud_nvws *ws, int64_t n, const double *x, const double *ewt, int64_t ng, double *y, double *acor,
                          double *del) {
    return run_reduce<R_SUM, 1>(ws, n, TNewtonUpdate{x, ewt, y, acor}, 1, 1, (double)(ng > 0 ? ng : n), del);
}
int shud_nv_dq_perturb(shud_nvws *ws, int64_t n, double sigma, const double *vs, const double *ewt, const double *y,
                       double *yt) {
    return run_map(ws, n, FDqPerturb{sigma, vs, ewt, y, yt});
}
int shud_nv_dq_combine(shud_nvws *ws, int64_t n, double sigma, double gamma, const double *vs, const double *ewt,
                       const double *fp, const double *fy, double *out) {
    return run_map(ws, n, FDqCombine{sigma, gamma, vs, ewt, fp, fy, out});
}
int shud_nv_wrmsnormvectorarray(shud_nvws *ws, int64_t n, int nv, const double *const *X, const double *const *W,
                                int64_t ng, double *out) {
    TWSqrMulti f;
    if (!fill(f.X, X, nv) || !fill(f.W, W, nv)) return SHUD_ERR_ARG;
    return run_reduce<R_SUM, SHUD_NV_MAXVEC>(ws, n, f, nv, 1, (double)(ng > 0 ? ng : n), out);
}

}  // extern "C"

// =============================================================================================
// SPGMR on the device: the linear solver CVLS hands every Newton iteration to
// (reference: SUNLinSol_SPGMR(udata, PREC_NONE, 0) + CVodeSetLinearSolver(mem, LS, NULL),
// src/Equations/cvode_config.cpp:172-179: maxl = 5, modified Gram-Schmidt, no restarts, matrix-free with
// difference-quotient J v).  The whole Arnoldi loop stays on the device: Gram-Schmidt coefficients are
// device scalars consumed by the next kernel, and the host synchronises ONCE per Krylov iteration to
// apply the Givens rotations and test convergence (k+2 scalars), instead of once per dot product.
// =============================================================================================
namespace {
struct FProdTo { const double *a, *b; double *z; __device__ void operator()(int64_t i) const { z[i] = a[i] * b[i]; } };
// fused passes of the modified Gram-Schmidt sweep (same arithmetic and the same reduction tree as the separate
// kernels: results are bit-identical; one read of w and one launch less per Krylov basis vector)
struct TDqCombineDot {  // out = S (I - gamma J) S^-1 v (FDqCombine);  term = out * u
    double sigma, gamma; const double *vs, *ewt, *fp, *fy; double *out; const double *u;
    __device__ double term(int, int64_t i) const {
        const double w = ewt[i], v = vs[i] / w;
        const double jv = (1.0 / sigma) * fp[i] + (-1.0 / sigma) * fy[i];
        const double o = w * (v + (-gamma) * jv);
        out[i] = o;
        return o * u[i];
    }
};
struct TAxpyNegDot {  // w -= h[0] * v;  term = w_new * u
    const double *h, *v; double *w; const double *u;
    __device__ double term(int, int64_t i) const {
        const double wn = w[i] + (-h[0]) * v[i];
        w[i] = wn;
        return wn * u[i];
    }
};
struct TAxpyNegSq {  // w -= h[0] * v;  term = w_new^2
    const double *h, *v; double *w;
    __device__ double term(int, int64_t i) const {
        const double wn = w[i] + (-h[0]) * v[i];
        w[i] = wn;
        return wn * wn;
    }
};
struct FNormalizeDev {  // w /= sqrt(n2[0])  (left untouched when the norm is 0)
    const double *n2; double *w;
    __device__ void operator()(int64_t i) const {
        const double s = sqrt(n2[0]);
        if (s != 0.0) w[i] = (1.0 / s) * w[i];
    }
};
struct FScaleTo { double c; const double *x; double *z; __device__ void operator()(int64_t i) const { z[i] = c * x[i]; } };
template <int NV>
struct FLinCombDiv {  // x = (sum c_k V_k) ./ ewt
    Coef c; Ptrs V; const double *ewt; double *x;
    __device__ void operator()(int64_t i) const { x[i] = lincomb_at<NV>(c, V, i) / ewt[i]; }
};
// Right-hand side of a Newton iteration, scaled, and its squared 2-norm in one pass:
// b = -((rl1 zn1 + acor) + (-gamma) f) (cvNlsResidual + the sign change of the Newton solver), V0 = ewt b (the first
// Krylov vector before normalisation); term = V0^2.  Same arithmetic, element by element, as N_VLinearCombination ->
// N_VScale(-1) -> N_VProd -> N_VDotProd; b itself is never stored.
struct TNewtonRhs {
    double rl1, gamma; const double *zn1, *acor, *f, *ewt; double *v0;
    __device__ double term(int, int64_t i) const {
        double r = rl1 * zn1[i];
        r += 1.0 * acor[i];
        r += (-gamma) * f[i];
        const double o = ewt[i] * (-1.0 * r);
        v0[i] = o;
        return o * o;
    }
};
// End of a Newton iteration in one pass: x = (sum c_k V_k) ./ ewt (FLinCombDiv), acor += x, y = zn0 + acor;
// term = (x ewt)^2 (the WRMS norm of the correction); x itself is never stored.
template <int NV>
struct TNewtonFinish {
    Coef c; Ptrs V; const double *ewt, *zn0; double *acor, *y;
    __device__ double term(int, int64_t i) const {
        const double w = ewt[i], a0 = acor[i], z0 = zn0[i];
        const double s = lincomb_at<NV>(c, V, i);
        const double x = s / w;
        const double a = a0 + 1.0 * x;
        acor[i] = a;
        y[i] = 1.0 * z0 + 1.0 * a;
        const double t = x * w;
        return t * t;
    }
};
// cvCompleteStep's update of the Nordsieck array, the weights of the next step and the norm of its tolsf test, and
// (CV_ONE_STEP) the copy of the solution to the caller's vector in one pass: zn[j] = l[j] acor + zn[j]
// (FScaleAddMulti), w = 1 / (rtol |zn0| + atol) (FEwt), term = (zn0 w)^2 (TWSqr), yout = zn0.
template <int Q>
struct TCompleteStep {
    Coef l; const double *acor; MPtrs Z; double rtol, atol; double *ewt, *yout;
    __device__ double term(int, int64_t i) const {
        const double xi = acor[i];
        double z[Q + 1];
#pragma unroll
        for (int j = 0; j <= Q; j++) z[j] = Z.p[j][i];
#pragma unroll
        for (int j = 0; j <= Q; j++) { z[j] = l.c[j] * xi + z[j]; Z.p[j][i] = z[j]; }
        const double z0 = z[0], w = 1.0 / (rtol * fabs(z0) + atol);
        ewt[i] = w;
        if (yout) yout[i] = 1.0 * z0;
        const double t = z0 * w;
        return t * t;
    }
};
// cvPredict / cvRestore on the whole Nordsieck array in one pass: the in-place Pascal-triangle sums
// zn[j-1] += sgn zn[j] (k = 1..q, j = q..k) run on registers, element by element in the order of the N_VLinearSum
// calls; optionally the start of the Newton iteration as well: acor = 0, y = zn[0] + acor.
template <int Q>
struct FPredict {
    double sgn; MPtrs Z; double *y, *acor;
    __device__ void operator()(int64_t i) const {
        double z[Q + 1];
#pragma unroll
        for (int j = 0; j <= Q; j++) z[j] = Z.p[j][i];
#pragma unroll
        for (int k = 1; k <= Q; k++)
#pragma unroll
            for (int j = Q; j >= k; j--) z[j - 1] = z[j - 1] + sgn * z[j];
#pragma unroll
        for (int j = 0; j < Q; j++) Z.p[j][i] = z[j];
        if (acor) { acor[i] = 0.0; y[i] = 1.0 * z[0] + 1.0 * 0.0; }
    }
};
}  // namespace

struct shud_spgmr {
    shud_ctx *gpu;
    shud_nvws *ws;
    int maxl;
    int64_t n;
    double sqrtN;
    std::vector<double *> V;   // maxl+1 Krylov vectors
    double *ytemp, *ftemp;
    double *dH;                // device scalars: Gram-Schmidt coefficients of the current column + squared norm
    double *hH;                // pinned host copy (mapped: the reductions of a single-GPU solve write it directly)
    double *hH_dev;            // its device alias
    int fold_dq = 1;           // shud_b200_rhs_dq_dev where the context allows it (SHUD_FOLD_DQ=0: separate perturbation)
    double *wraw = nullptr;    // folded route: the unnormalised Krylov vector in the making (r0, then each w); the RHS
                               // pre-pass of the next iteration normalises it into V[k] on the way
};

// the folded route: a single domain whose RHS pre-pass perturbs and normalises (shud_b200_rhs_dq_dev)
static bool spgmr_folds(const shud_spgmr *s) {
    const shud_nvws *ws = s->ws;
    return s->fold_dq && s->wraw && !(ws->ar_dev && !ws->ar_off) && shud_b200_dq_foldable(s->gpu);
}

// a sum over the (distributed) vector into dH[0] and hH[0]
template <class F>
static int reduce_to_host(shud_spgmr *s, F f) {
    shud_nvws *ws = s->ws;
    int rc;
    if (ws->ar_dev && !ws->ar_off && ws->peer.nranks <= 1) {
        if ((rc = run_reduce_dev<R_SUM>(ws, s->n, f, s->dH))) return rc;
        CKN(cudaMemcpyAsync(s->hH, s->dH, sizeof(double), cudaMemcpyDeviceToHost, ws->stream));
        CKN(cudaStreamSynchronize(ws->stream));
        return SHUD_OK;
    }
    const double ticket = (ws->ticket += 1.0);
    if ((rc = run_reduce_dev<R_SUM>(ws, s->n, f, s->dH, s->hH_dev, ticket))) return rc;
    return wait_ticket(ws, ticket);
}


extern "C" {

int shud_spgmr_create(shud_ctx *gpu, shud_nvws *ws, int maxl, int64_t n_global, shud_spgmr **out) {
    if (!gpu || !ws || !out || maxl < 1 || maxl >= SHUD_NV_MAXVEC) return SHUD_ERR_ARG;
    if ((void *)ws->stream != shud_b200_stream(gpu)) return SHUD_ERR_ARG;  // RHS and vector work share one stream
    shud_spgmr *s = new shud_spgmr();
    s->gpu = gpu; s->ws = ws; s->maxl = maxl; s->n = shud_b200_ny(gpu);
    if (const char *e = getenv("SHUD_FOLD_DQ")) s->fold_dq = atoi(e);
    s->sqrtN = sqrt((double)(n_global > 0 ? n_global : s->n));
    CKN(cudaSetDevice(ws->device));
    for (int k = 0; k <= maxl; k++) {
        double *p = nullptr;
        CKN(cudaMalloc(&p, sizeof(double) * s->n));
        s->V.push_back(p);
    }
    CKN(cudaMalloc(&s->wraw, sizeof(double) * s->n));
    CKN(cudaMalloc(&s->ytemp, sizeof(double) * s->n));
    CKN(cudaMalloc(&s->ftemp, sizeof(double) * s->n));
    CKN(cudaMalloc(&s->dH, sizeof(double) * (maxl + 2)));
    CKN(cudaHostAlloc(&s->hH, sizeof(double) * (maxl + 2), cudaHostAllocMapped));
    CKN(cudaHostGetDevicePointer((void **)&s->hH_dev, s->hH, 0));
    *out = s;
    return SHUD_OK;
}

void shud_spgmr_set_nglobal(shud_spgmr *s, int64_t n_global) {
    if (s && n_global > 0) s->sqrtN = sqrt((double)n_global);
}

void shud_spgmr_destroy(shud_spgmr *s) {
    if (!s) return;
    cudaSetDevice(s->ws->device);
    cudaDeviceSynchronize();
    for (double *p : s->V) cudaFree(p);
    cudaFree(s->wraw); cudaFree(s->ytemp); cudaFree(s->ftemp); cudaFree(s->dH); cudaFreeHost(s->hH);
    delete s;
}

// Arnoldi / modified Gram-Schmidt iterations from the scaled residual in V[0] (2-norm beta): normalises V[0], runs
// up to maxl iterations (one RHS call each), solves the small least-squares problem; yk[0..k_used) are the
// coefficients of the correction in the Krylov basis.
static int spgmr_iterate(shud_spgmr *s, double t, double gamma, const double *y, const double *fy, const double *ewt,
                         double beta, double tol, double *yk, int *k_used_out, double *res_out, bool *conv_out) {
    shud_nvws *ws = s->ws;
    const int64_t n = s->n;
    const int maxl = s->maxl;
    int rc;
    // folded route (single domain): r0 lies unnormalised in wraw with its squared norm in dH[0]; the pre-pass of each
    // RHS call normalises the vector of the iteration into V[k] while it perturbs with it
    const bool fold = spgmr_folds(s);
    if (!fold && (rc = run_map(ws, n, FScaleTo{1.0 / beta, s->V[0], s->V[0]}))) return rc;
    double H[SHUD_NV_MAXVEC + 1][SHUD_NV_MAXVEC] = {{0}};
    double g[SHUD_NV_MAXVEC + 1] = {0}, cs[SHUD_NV_MAXVEC] = {0}, sn[SHUD_NV_MAXVEC] = {0};
    g[0] = beta;
    int k_used = 0;
    bool conv = false;
    const double sig = s->sqrtN;  // 1/||S^-1 v_k||_WRMS: v_k has unit 2-norm (CVLS' sigma without a reduction)
    const bool dist = ws->ar_dev && !ws->ar_off;  // a partition of a multi-GPU run: f() = halo exchange + RHS
    for (int k = 0; k < maxl; k++) {
        // w = S (I - gamma J) S^-1 v_k, J by difference quotient: 1 RHS call
        // (single domain: the perturbation is formed by the RHS's own pre-pass)
        if (fold) {
            if ((rc = shud_b200_rhs_dq_dev(s->gpu, t, sig, s->wraw, ewt, y, s->ytemp, s->ftemp, s->dH + k, s->V[k]))) return rc;
        } else {
            if ((rc = shud_nv_dq_perturb(ws, n, sig, s->V[k], ewt, y, s->ytemp))) return rc;
            if ((rc = dist ? shud_b200_rhs_exchange_dev(s->gpu, t, s->ytemp, s->ftemp) : shud_b200_rhs_dev(s->gpu, t, s->ytemp, s->ftemp))) return rc;
        }
        // ... fused with the first dot product of the modified Gram-Schmidt sweep; every later pass subtracts the
        // previous projection and forms the next dot product (the last one the squared norm) in one read of w.
        // Coefficients stay on the device: h_i is written by the reduction's last block and read by the next pass.
        double *w = fold ? s->wraw : s->V[k + 1];
        // one GPU: every scalar also lands in mapped host memory, the last one with a ticket the host spins on (the
        // normalisation below runs while the host does the Givens rotations); distributed: the allreduced scalars
        // are copied back and the stream synchronised
        const bool via_copy = dist && ws->peer.nranks <= 1;  // scalars allreduced by NCCL behind the kernels
        double *const hd = via_copy ? nullptr : s->hH_dev;
        const double ticket = via_copy ? 0.0 : (ws->ticket += 1.0);
        if ((rc = run_reduce_dev<R_SUM>(ws, n, TDqCombineDot{sig, gamma, s->V[k], ewt, s->ftemp, fy, w, s->V[0]}, s->dH, hd))) return rc;
        for (int i = 0; i < k; i++)
            if ((rc = run_reduce_dev<R_SUM>(ws, n, TAxpyNegDot{s->dH + i, s->V[i], w, s->V[i + 1]}, s->dH + i + 1, hd ? hd + i + 1 : nullptr))) return rc;
        if ((rc = run_reduce_dev<R_SUM>(ws, n, TAxpyNegSq{s->dH + k, s->V[k], w}, s->dH + k + 1, hd ? hd + k + 1 : nullptr, ticket))) return rc;
        if (!fold && (rc = run_map(ws, n, FNormalizeDev{s->dH + k + 1, s->V[k + 1]}))) return rc;
        if (via_copy) {
            CKN(cudaMemcpyAsync(s->hH, s->dH, sizeof(double) * (k + 2), cudaMemcpyDeviceToHost, ws->stream));
            CKN(cudaStreamSynchronize(ws->stream));  // the one host synchronisation of this Krylov iteration
        } else if ((rc = wait_ticket(ws, ticket))) return rc;
        for (int i = 0; i <= k; i++) H[i][k] = s->hH[i];
        H[k + 1][k] = sqrt(s->hH[k + 1]);
        for (int i = 0; i < k; i++) {
            const double tmp = cs[i] * H[i][k] + sn[i] * H[i + 1][k];
            H[i + 1][k] = -sn[i] * H[i][k] + cs[i] * H[i + 1][k];
            H[i][k] = tmp;
        }
        const double den = hypot(H[k][k], H[k + 1][k]);
        if (den != 0.0) { cs[k] = H[k][k] / den; sn[k] = H[k + 1][k] / den; } else { cs[k] = 1.0; sn[k] = 0.0; }
        H[k][k] = cs[k] * H[k][k] + sn[k] * H[k + 1][k];
        H[k + 1][k] = 0.0;
        g[k + 1] = -sn[k] * g[k];
        g[k] = cs[k] * g[k];
        k_used = k + 1;
        if (fabs(g[k + 1]) <= tol) { conv = true; break; }
    }
    for (int i = 0; i < SHUD_NV_MAXVEC; i++) yk[i] = 0.0;
    for (int i = k_used - 1; i >= 0; i--) {
        double acc = g[i];
        for (int j = i + 1; j < k_used; j++) acc -= H[i][j] * yk[j];
        yk[i] = acc / H[i][i];
    }
    *k_used_out = k_used;
    *res_out = fabs(g[k_used]);
    *conv_out = conv;
    return 0;
}

int shud_spgmr_solve(shud_spgmr *s, double t, double gamma, const double *y, const double *fy, const double *ewt,
                     const double *b, double tol, double *x, int *nli_out, double *res_out) {
    if (!s || !y || !fy || !ewt || !b || !x) return SHUD_ERR_ARG;
    shud_nvws *ws = s->ws;
    const int64_t n = s->n;
    int rc;
    // r0 = S b, beta = ||r0||_2
    double *const r0 = spgmr_folds(s) ? s->wraw : s->V[0];
    if ((rc = run_map(ws, n, FProdTo{ewt, b, r0}))) return rc;
    if ((rc = reduce_to_host(s, TDot{r0, r0}))) return rc;
    const double beta = sqrt(s->hH[0]);
    if (nli_out) *nli_out = 0;
    if (res_out) *res_out = beta;
    if (beta <= tol) {
        if ((rc = run_map(ws, n, FConst{0.0, x}))) return rc;
        return 0;
    }
    double yk[SHUD_NV_MAXVEC], res = 0.0;
    int k_used = 0;
    bool conv = false;
    if ((rc = spgmr_iterate(s, t, gamma, y, fy, ewt, beta, tol, yk, &k_used, &res, &conv))) return rc;
    Coef cf; Ptrs P;
    for (int i = 0; i < k_used; i++) { cf.c[i] = yk[i]; P.p[i] = s->V[i]; }
    switch (k_used) {
        case 1: rc = run_map(ws, n, FLinCombDiv<1>{cf, P, ewt, x}); break;
        case 2: rc = run_map(ws, n, FLinCombDiv<2>{cf, P, ewt, x}); break;
        case 3: rc = run_map(ws, n, FLinCombDiv<3>{cf, P, ewt, x}); break;
        case 4: rc = run_map(ws, n, FLinCombDiv<4>{cf, P, ewt, x}); break;
        case 5: rc = run_map(ws, n, FLinCombDiv<5>{cf, P, ewt, x}); break;
        case 6: rc = run_map(ws, n, FLinCombDiv<6>{cf, P, ewt, x}); break;
        default: rc = run_map(ws, n, FLinCombDiv<7>{cf, P, ewt, x}); break;
    }
    if (rc) return rc;
    if (nli_out) *nli_out = k_used;
    if (res_out) *res_out = res;
    return conv ? 0 : (res < beta ? 1 : 2);
}

int shud_spgmr_newton_step(shud_spgmr *s, double t, double gamma, double rl1, const double *zn0, const double *zn1,
                           double *acor, double *y, const double *fy, const double *ewt, double tol, int64_t n_global,
                           double *del, int *nli_out, double *res_out) {
    if (!s || !zn0 || !zn1 || !acor || !y || !fy || !ewt || !del) return SHUD_ERR_ARG;
    shud_nvws *ws = s->ws;
    const int64_t n = s->n;
    int rc;
    if ((rc = reduce_to_host(s, TNewtonRhs{rl1, gamma, zn1, acor, fy, ewt, spgmr_folds(s) ? s->wraw : s->V[0]}))) return rc;
    const double beta = sqrt(s->hH[0]);
    if (nli_out) *nli_out = 0;
    if (res_out) *res_out = beta;
    if (beta <= tol) return 3;  // nothing written: the caller takes the unfused route for this iteration
    double yk[SHUD_NV_MAXVEC], res = 0.0;
    int k_used = 0;
    bool conv = false;
    if ((rc = spgmr_iterate(s, t, gamma, y, fy, ewt, beta, tol, yk, &k_used, &res, &conv))) return rc;
    Coef cf; Ptrs P;
    for (int i = 0; i < k_used; i++) { cf.c[i] = yk[i]; P.p[i] = s->V[i]; }
    const double ng = (double)(n_global > 0 ? n_global : n);
    switch (k_used) {
        case 1: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<1>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
        case 2: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<2>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
        case 3: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<3>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
        case 4: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<4>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
        case 5: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<5>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
        case 6: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<6>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
        default: rc = run_reduce<R_SUM, 1>(ws, n, TNewtonFinish<7>{cf, P, ewt, zn0, acor, y}, 1, 1, ng, del); break;
    }
    if (rc) return rc;
    if (nli_out) *nli_out = k_used;
    if (res_out) *res_out = res;
    return conv ? 0 : (res < beta ? 1 : 2);
}

int shud_nv_bdf_complete(shud_nvws *ws, int64_t n, int q, const double *l, const double *acor, double *const *zn, double rtol,
                         double atol, double *ewt, double *yout, int64_t ng, double *nrm) {
    if (!l || !acor || !zn || !ewt || !nrm || q < 1 || q > 5) return SHUD_ERR_ARG;
    MPtrs Z; Coef c;
    if (!fill(Z, zn, q + 1)) return SHUD_ERR_ARG;
    for (int j = 0; j <= q; j++) c.c[j] = l[j];
    const double g = (double)(ng > 0 ? ng : n);
    switch (q) {
        case 1: return run_reduce<R_SUM, 1>(ws, n, TCompleteStep<1>{c, acor, Z, rtol, atol, ewt, yout}, 1, 1, g, nrm);
        case 2: return run_reduce<R_SUM, 1>(ws, n, TCompleteStep<2>{c, acor, Z, rtol, atol, ewt, yout}, 1, 1, g, nrm);
        case 3: return run_reduce<R_SUM, 1>(ws, n, TCompleteStep<3>{c, acor, Z, rtol, atol, ewt, yout}, 1, 1, g, nrm);
        case 4: return run_reduce<R_SUM, 1>(ws, n, TCompleteStep<4>{c, acor, Z, rtol, atol, ewt, yout}, 1, 1, g, nrm);
        default: return run_reduce<R_SUM, 1>(ws, n, TCompleteStep<5>{c, acor, Z, rtol, atol, ewt, yout}, 1, 1, g, nrm);
    }
}

int shud_nv_bdf_predict(shud_nvws *ws, int64_t n, int q, double sgn, double *const *zn, double *y, double *acor) {
    if (!zn || q < 1 || q > 5 || (acor && !y)) return SHUD_ERR_ARG;
    MPtrs Z;
    if (!fill(Z, zn, q + 1)) return SHUD_ERR_ARG;
    switch (q) {
        case 1: return run_map(ws, n, FPredict<1>{sgn, Z, y, acor});
        case 2: return run_map(ws, n, FPredict<2>{sgn, Z, y, acor});
        case 3: return run_map(ws, n, FPredict<3>{sgn, Z, y, acor});
        case 4: return run_map(ws, n, FPredict<4>{sgn, Z, y, acor});
        default: return run_map(ws, n, FPredict<5>{sgn, Z, y, acor});
    }
}

}  // extern "C"
