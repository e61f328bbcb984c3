// shud_rhs.cu - the SHUD right-hand side f(t,y,ydot) on one B200 (sm_100a), FP64, SoA.
//
// Replaces the reference's f() = f_update + f_loop + f_applyDY
// (src/Model/f.cpp:2-32, src/ModelData/MD_update.cpp:102-189, MD_f.cpp:9-215) behind the C ABI
// of include/shud_b200.h.  Layout and kernels are described in DESIGN.md; in short:
//   * cells are renumbered along a Hilbert curve (locality of the 3 neighbour gathers), reaches
//     follow the cells their segments touch, segments are grouped by reach;
//   * every flux is evaluated owner-computes (cell i computes its own 3 edges and its own
//     river segments; a reach re-evaluates its upstream reaches' Manning flux) - no atomics,
//     every sum runs in a fixed order, so ydot is run-to-run reproducible;
//   * launches per RHS: effKH pre-pass, warp-specialised fused cell kernel, river+lake kernel.
// Device vectors are in DEVICE ORDER (permuted); shud_b200_rhs() (host pointers, reference
// order) permutes on the way in and out.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "shud_b200.h"
#include "shud_nvector.h"  // SHUD_NV_ARBOX_BYTES, SHUD_NV_MAXRANKS: the mailbox region of the in-kernel allreduce
#include "shud_phys.cuh"

namespace {

using namespace shud;

constexpr unsigned F_LAKE = 1u, F_HEADBC = 2u, F_FLUXBC = 4u, F_SS_SURF = 8u, F_SS_GW = 16u, F_GHOST = 32u;
constexpr int NSEG_SHIFT = 8;

struct DevMesh {
    int Ne, Nr, Ns, Nl, close_boundary, has_headbc, nbank;
    int ld;  // Ne rounded up to a multiple of 128: every per-cell array this library allocates is padded to it
    // static per cell
    const double *area, *z_surf, *z_bottom, *depression, *aqd, *sy, *infD, *infKsatV, *macKsatV, *hAreaF, *thetaS,
        *thetaR, *thetaFC, *beta, *ksatH, *ksatV, *macKsatH, *macD, *vAreaF, *vegFrac, *impAF, *wetland, *rootReach,
        *rough, *qss;
    const double *edge, *dist, *dist2edge, *avgRough;  // [3][Ne]
    const int *nbr;                                    // [3][Ne]: >=0 cell, -1 boundary, <=-2 bank slot (-2-v)
    const unsigned *flags;
    const int *cell_seg_first;  // first slot of the cell in the cell-ordered segment table below
    // cell-ordered segment table (slot = cell_seg_first[i] + k, ascending reference segment id per cell):
    // everything static the cell side of a segment needs, copied next to each other so that the only
    // dependent gather left is the reach stage
    const int *cs_seg, *cs_riv, *cs_bc;  // segment id (device order), reach id (device order), reach BC code
    const int *cs_cell;                  // owning cell (device order) of the slot
    const double *cs_len, *cs_cwr, *cs_depth, *cs_zbank, *cs_ksatH, *cs_bed;
    const double *cs_zr, *cs_zbk;        // z_surf[cell] - depth (river bed), z_surf[cell] + zbank (bank top): static
    double *cs_yr;                       // stage of the slot's reach for this call (BC applied), written by k_effkh
    // forcing step
    double *netPrep, *potEvap, *potTran, *lai, *fuSurf, *fuSub, *eic, *satn, *ele_yBC, *ele_QBC;
    // work
    double *effKH, *QsegSurf, *QsegSub;
    // reaches
    const double *r_len, *r_slope, *r_depth, *r_w0, *r_bank, *r_rough, *r_dist, *r_ksatH, *r_bed, *r_zbank;
    const int *r_down, *r_bc, *r_toLake, *r_up_ptr, *r_up_idx, *r_seg_ptr;
    double *r_yBC, *r_qBC;
    // segments (device order = grouped by reach)
    const int *s_riv;
    const double *s_len, *s_cwr;
    // lakes
    const double *l_zmin, *l_yi0, *l_by, *l_ba;
    const int *l_bptr, *l_bank_ptr, *l_rin_ptr, *l_rin_idx;
    const int *bank_cell, *bank_j, *bank_lake;
    const double *bank_kh;
    double *l_evap_raw, *l_prcp;
    // halo cells of a partition (ids Ne .. Ne+Nhalo-1 in nbr)
    int Nhalo;
    const double *h_zs, *h_zb, *h_aqd, *h_macD, *h_macKsatH, *h_vAreaF, *h_ksatH;
    const double *h_state;        // [Nhalo][2] = (Ysurf, Ygw) of each halo cell, filled by the halo exchange
    // peer-to-peer exchange: two halo buffers written alternately by the neighbours (h_state = even, h_state_alt = odd
    // epochs); the epoch of the exchange in flight is a device word (NULL: h_state as registered)
    const double *h_state_alt;
    unsigned long long *h_epoch;           // completed exchanges; the one in flight is *h_epoch + 1
    unsigned int *h_done;                  // blocks of the river / lake kernel through (the last one advances the epoch)
    const unsigned long long *h_flags;     // [h_nflags] epoch of the last halo each neighbour has delivered here
    int h_nflags, n_int_tiles;             // tiles >= n_int_tiles see halo cells: they wait for the flags
    // cut river trees: ghost cells / ghost reaches (owned elsewhere, evaluated here from exchanged states; ydot 0)
    const int *g_cslot;                    // [Ne] ghost-cell slot of a cell flagged F_GHOST
    const int *r_gslot;                    // [Nr] ghost-reach slot, -1 = own reach (NULL: no ghost reach)
    const int *cs_g;                       // [Ns] per cell-side segment slot: ghost-reach slot of its reach or -1
    int g_coff, g_roff;                    // where the ghost-cell triples / ghost-reach stages start in a halo buffer
    int *err;  // [0] code, [1] where (1-based reference id)
};

struct DevDiag {
    double *qEleInfil, *qEleExfil, *qEleRecharge, *qEs, *qEu, *qEg, *qTu, *qTg, *qEleTrans, *qEleEvapo, *qEleETA,
        *iBeta, *QeleSurf, *QeleSub, *QeleSurfTot, *QeleSubTot, *Qe2r_Surf, *Qe2r_Sub, *QrivSurf, *QrivSub, *QrivUp,
        *QrivDown, *y2LakeArea, *QLakeSurf, *QLakeSub, *QLakeRivIn, *QLakeRivOut, *qLakeEvap, *qLakePrcp;
};

// land-surface step (shud_land.cuh): statics, bucket states, per-step tables, outputs kept for the host
struct DevLand {
    int nforc, nlc, nmf;
    const int *iForc, *iLC, *iMF;
    const double *albedo, *fixP, *windH, *nx, *ny, *nz, *forc_z;
    double cPrep, cTemp, cLAItsd, cMF, cETP, cISmax;
    int net, tsr;
    double cap, cosz_min;
    double *snow, *ics;                         // yEleSnow, yEleIS (device order)
    double *cls;                                // per-step, per land-cover class: lai | soil-heat factor | log log | rs (k_land_tables)
    double *tab;                                // per-step tables: forc[5 nforc] | lai[nlc] | mf[nmf] | sx|sy|sz|wdt [tsr_cap each]
    int tsr_cap;
    double *prep, *etp, *temp, *tmf, *factor;   // qElePrep, qEleETP, t_temp, t_mf, terrain factor
    const int *lk_ptr, *lk_cell;                // lake -> its cells (device ids), ascending reference id
    const double *lk_rnele;                     // lake -> (double)NumEleLake
    // CRYOSPHERE = 1: the two _AccTemp accumulators of every cell (AccTemperature.hpp): day sum, running sums,
    // rings of the last Ls / Lb daily means [slot][ld]
    int cryo, Ls, Lb;
    double surf_max, surf_min, sub_max, sub_min;
    double *tacc, *acc_s, *acc_b, *ring_s, *ring_b;
};
struct CryoStep {  // per-step, uniform over the cells (the day clock of the accumulators lives on the host)
    int do_push, pop_s, pop_b, slot_s, slot_b;
    double nday, size_s, size_b;
    double r_nday, r_size_s, r_size_b;  // their reciprocals (SHUD_RCP)
};

// ---------------------------------------------------------------------------------------------
// Peer-to-peer halo exchange over NVLink (one process per GPU, the neighbours' halo buffers mapped through CUDA IPC),
// fused into the RHS launches: the first blocks of the pre-pass store each boundary cell's (Ysurf, Ygw) straight into
// the halo buffer of the partition that needs it and, when the last of them is through, release one flag per
// neighbour; the tiles of the cell kernel that see halo cells are ordered last in the grid and acquire the flags
// before their first halo read (by then, ~100 us later, the flags have long arrived).  No collective call, no staging
// buffer, no extra launch, no second stream: a partition's f() is the same three launches as a single domain's.
// Epochs: the exchange of call e writes buffer e & 1.  A neighbour can be at most one call ahead (it cannot pack call
// e + 2 before it has seen my flag of call e + 1, which I release after everything of call e has completed), so two
// buffers suffice and no credit has to travel back.  The epoch word advances at the end of the river / lake kernel.
// ---------------------------------------------------------------------------------------------
constexpr int P2P_MAXPEER = 16;
constexpr int P2P_MAXSEG = 3 * P2P_MAXPEER;
struct P2PTable {
    double *buf[2][P2P_MAXPEER];            // neighbour p's halo buffers (even / odd epochs), peer-mapped
    unsigned long long *flag[P2P_MAXPEER];  // my flag slot in neighbour p's flag array
    int npeers, nseg;
    // the doubles I send are grouped by (neighbour, kind): segment s = items [seg_start[s], seg_start[s + 1]) go to
    // neighbour seg_peer[s], starting at double seg_dst[s] of its halo buffer
    int seg_start[P2P_MAXSEG + 1], seg_peer[P2P_MAXSEG], seg_dst[P2P_MAXSEG];
};
struct PackArgs {
    P2PTable T;
    const int *idx;        // flat device-order indices into the state vector of the doubles sent
    int n, nblk;           // doubles sent, blocks of the pre-pass that carry them
    unsigned int *count;   // blocks through
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ const double *halo_state(const DevMesh &m) {
    if (m.h_epoch) return ((*m.h_epoch + 1ull) & 1ull) ? m.h_state_alt : m.h_state;
    return m.h_state;
}

__device__ __forceinline__ void raise_err(int *err, int code, int where) {
    if (atomicMax_system(&err[0], code) < code) err[1] = where;
}

// ---------------------------------------------------------------------------------------------
// K0: horizontal effective conductivity of every cell (neighbours need it before any edge flux)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void effkh_body(const DevMesh &m, const double *__restrict__ Y) {
    {
        // stage of every segment slot's reach, so that the cell kernel reads it without a dependent gather
        const size_t NE3 = 3 * (size_t)m.Ne;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < m.Ns; q += gridDim.x * blockDim.x) {
            const int r = __ldg(m.cs_riv + q);
            m.cs_yr[q] = (__ldg(m.cs_bc + q) > 0) ? m.r_yBC[r] : Y[NE3 + r];
        }
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.Ne) return;
    const unsigned fl = m.flags[i];
    double kh;
    if (fl & F_LAKE) {
        kh = m.ksatH[i];  // _Element::updateLakeElement, src/classes/Element.cpp:336-337
    } else {
        const double ygw = (fl & F_HEADBC) ? m.ele_yBC[i] : Y[2 * (size_t)m.Ne + i];
        int e = 0;
        kh = eff_kh(ygw, m.aqd[i], m.macD[i], m.macKsatH[i], m.vAreaF[i], m.ksatH[i], &e);
        if (e) raise_err(m.err, e, i + 1);
    }
    m.effKH[i] = kh;
}
__global__ void __launch_bounds__(256) k_effkh(DevMesh m, const double *__restrict__ Y) {
    // programmatic dependent launch: the cell kernel may start now; it waits (griddepcontrol.wait) only where it
    // first needs effKH, so its vertical role overlaps this pre-pass
    asm volatile("griddepcontrol.launch_dependents;");
    effkh_body(m, Y);
}
// The pre-pass of a difference-quotient evaluation f(y0 + sigma v / ewt) (CVLS' Jv inside SPGMR): it forms the
// perturbed state - the arithmetic of shud_nv_dq_perturb, element by element - stores it for the cell and river
// kernels, and evaluates effKH from the value it has just formed: one pass over (v, ewt, y0) instead of a vector
// kernel followed by a pre-pass that reads the result back.  A reach stage another thread forms is re-formed here.
// With ss != NULL the direction is given unnormalised: v[k] / sqrt(*ss) is what is perturbed with - the arithmetic of the
// Krylov solver's normalisation kernel, (1 / sqrt(ss)) * v - and the normalised direction is stored to v_out (a buffer
// other than v: another thread may still re-form a reach stage from v).  It saves the solver a pass per Krylov vector.
struct DqArgs { double sigma; const double *v, *ewt, *y0; double *yt; const double *ss; double *v_out; };
__device__ __forceinline__ double dq_dir(const DqArgs &A, size_t k, double rs) { return A.ss ? rs * A.v[k] : A.v[k]; }
__device__ __forceinline__ double dq_entry(const DqArgs &A, size_t k, double rs) {  // an entry this thread owns
    const double v = dq_dir(A, k, rs);
    if (A.v_out) A.v_out[k] = v;
    const double yt = A.sigma * (v / A.ewt[k]) + A.y0[k];
    A.yt[k] = yt;
    return yt;
}
__global__ void __launch_bounds__(256) k_effkh_dq(DevMesh m, DqArgs A) {
    const size_t NE = (size_t)m.Ne, NE3 = 3 * NE;
    double rs = 1.0;
    if (A.ss) { const double s = sqrt(*A.ss); rs = s != 0.0 ? 1.0 / s : 1.0; }
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < m.Ns; q += gridDim.x * blockDim.x) {
        const int r = __ldg(m.cs_riv + q);
        m.cs_yr[q] = (__ldg(m.cs_bc + q) > 0) ? m.r_yBC[r] : A.sigma * (dq_dir(A, NE3 + r, rs) / A.ewt[NE3 + r]) + A.y0[NE3 + r];
    }
    for (size_t k = NE3 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < NE3 + m.Nr + m.Nl; k += (size_t)gridDim.x * blockDim.x)
        dq_entry(A, k, rs);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.Ne) return;
    dq_entry(A, i, rs);
    dq_entry(A, NE + i, rs);
    const double ygw_t = dq_entry(A, 2 * NE + i, rs);
    const unsigned fl = m.flags[i];
    double kh;
    if (fl & F_LAKE) {
        kh = m.ksatH[i];
    } else {
        const double ygw = (fl & F_HEADBC) ? m.ele_yBC[i] : ygw_t;
        int e = 0;
        kh = eff_kh(ygw, m.aqd[i], m.macD[i], m.macKsatH[i], m.vAreaF[i], m.ksatH[i], &e);
        if (e) raise_err(m.err, e, i + 1);
    }
    m.effKH[i] = kh;
}
// the same with the send side of the peer-to-peer halo exchange in its first blocks
__global__ void __launch_bounds__(256) k_effkh_pack(DevMesh m, const double *__restrict__ Y, PackArgs P) {
    asm volatile("griddepcontrol.launch_dependents;");
    if ((int)blockIdx.x < P.nblk) {
        const unsigned long long e = *m.h_epoch + 1ull;  // the epoch word advances at the end of this call (k_river_lake)
        // (a small partition can have more doubles to send than pre-pass threads: the packing blocks stride over them)
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += P.nblk * blockDim.x) {
            int sg = 0;
            while (sg + 1 < P.T.nseg && k >= P.T.seg_start[sg + 1]) sg++;
            const int p = P.T.seg_peer[sg];
            double *dst = ((e & 1ull) ? P.T.buf[1][p] : P.T.buf[0][p]) + (size_t)(P.T.seg_dst[sg] + (k - P.T.seg_start[sg]));
            *dst = Y[P.idx[k]];  // a store over NVLink into the neighbour's halo buffer
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int done = atomicAdd(P.count, 1u);
            if (done == (unsigned)P.nblk - 1) {  // every packing block's stores are fenced: publish
                __threadfence();
                *P.count = 0;
                for (int p = 0; p < P.T.npeers; p++) st_release_sys(P.T.flag[p], e);
            }
        }
    }
    effkh_body(m, Y);
}

// The partition-only pieces of the cell kernel.  They are rare paths (flag acquire and ghost cells in a few tiles, halo
// neighbours on a few edges) but must stay inlined: as real calls (__noinline__) they impose the ABI's register
// discipline on the whole kernel - measured 405 us instead of 108.5 us for the partition instantiation.
#define SHUD_HALO_NOINLINE __forceinline__
// one lane per neighbour partition acquires that neighbour's flag of the exchange in flight
__device__ SHUD_HALO_NOINLINE void acquire_flag(const DevMesh &m, int lane) {
    const unsigned long long e = *m.h_epoch + 1ull;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(m.h_flags + lane) < e) {
        __nanosleep(200);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if (t1 - t0 > 5000000000ull) {  // 5 s: a neighbour is gone; report instead of hanging the device
            raise_err(m.err, SHUD_ERR_P2P_TIMEOUT, lane + 1);
            break;
        }
    }
}
// halo neighbour h: state from the last exchange, statics sent once, effKH evaluated here
__device__ SHUD_HALO_NOINLINE void halo_neighbour(const DevMesh &m, int h, double &nsf, double &ygw_n, double &zs_n,
                                                  double &zb_n, double &kh_n) {
    const double *hs = halo_state(m);
    nsf = hs[2 * h]; ygw_n = hs[2 * h + 1]; zs_n = m.h_zs[h]; zb_n = m.h_zb[h];
    int e2 = 0;  // a range violation is reported by the partition that owns the cell
    kh_n = eff_kh(ygw_n, m.h_aqd[h], m.h_macD[h], m.h_macKsatH[h], m.h_vAreaF[h], m.h_ksatH[h], &e2);
}
// ghost cell ic: (Ysurf, Yunsat, Ygw) from the exchange, and its effKH (the pre-pass saw its stale vector entry)
__device__ SHUD_HALO_NOINLINE void ghost_state(const DevMesh &m, int ic, unsigned fl, double &gsf, double &gus, double &ggw,
                                               double &gkh) {
    const double *g = halo_state(m) + m.g_coff + 3 * (size_t)m.g_cslot[ic];
    gsf = __ldcg(g); gus = __ldcg(g + 1); ggw = __ldcg(g + 2);
    const double head = (fl & F_HEADBC) ? m.ele_yBC[ic] : ggw;
    int e2 = 0;
    gkh = (fl & F_LAKE) ? m.ksatH[ic] : eff_kh(head, m.aqd[ic], m.macD[ic], m.macKsatH[ic], m.vAreaF[ic], m.ksatH[ic], &e2);
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised fused cell kernel.  A 256-thread block owns a tile of 128 consecutive cells
// (consecutive along the Hilbert curve, so a compact patch of the mesh), two threads per cell:
//   warps 0-3 ("lateral" role): every own value and edge static of the cell lands in shared memory by cp.async;
//       the tile's Ysurf, Ygw, z_surf, z_bottom, effKH double as the neighbour table (global memory only for
//       the ~15 % of neighbours outside the tile / in the halo); 3 overland + 3 groundwater edge fluxes (one
//       copy of the edge code, three trips); the tile's river segments, one lane per segment slot (groundwater
//       exchange before the hand-over, weir after it); Qe2r + Q0 + Q1 + Q2 and the first two terms of dYsf, dYgw;
//   warps 4-7 ("vertical" role): its 26 inputs by cp.async, then in steps that re-read what they need from the
//       staged slots (nothing held in a register across a pow()): updateElement -> infiltration / recharge ->
//       hand P1, G1 and the ponding left for the weir to the lateral role -> ET partition while the lateral role
//       finishes -> the three balance equations, ydot stores, carried state.
// 64 registers, 48 KB of shared memory, 4 blocks (32 warps) per SM, no local-memory spills in the hot path, no
// extra HBM traffic against the one-thread-per-cell form.  No atomics; every sum in a fixed order.
// Named barriers: 1 lateral-internal, 2 vertical -> lateral hand-over, 3 lateral -> vertical hand-back.
// Programmatic dependent launch: starts under k_effkh (the lateral role waits before its effKH copy) and lets
// k_river_lake start in its last wave.
// ---------------------------------------------------------------------------------------------
constexpr int TILE = 128;
constexpr int V_NIN = 26;     // per-cell inputs of the vertical role
constexpr int SEGCAP = 256;  // segment slots per tile kept in shared memory (more: read back from global)
// 8-byte asynchronous copy global -> shared (SASS: LDGSTS)
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// HALO: the context is a partition (halo cells, flag acquire of the peer-to-peer exchange); a single domain compiles
// neither (the cell kernel is sensitive to its instruction footprint: 106.4 vs 102.6 us with the halo code present)
template <bool DIAG, int MINB, int HALO>
__global__ void __launch_bounds__(2 * TILE, MINB) k_fused(DevMesh m, DevDiag d, const double *__restrict__ Y,
                                                          double *__restrict__ DY, int tile0) {
    __shared__ double t_sf[TILE], t_gw[TILE], t_zs[TILE], t_zb[TILE], t_kh[TILE], t_dep[TILE], t_fus[TILE], t_area[TILE];
    __shared__ double e_B[3][TILE], e_dist[3][TILE], e_rough[3][TILE];  // per-edge statics of the lateral role
    __shared__ int t_seg0[TILE];
    __shared__ unsigned t_flv[TILE];  // the vertical role's copy of the cell flags (not worth a register for 3 steps)
    __shared__ double sq_s[SEGCAP], sq_g[SEGCAP];  // river-segment fluxes of the tile, slot order
    __shared__ double v_in[V_NIN][TILE];           // inputs of the vertical role, landed by cp.async
    // values handed between the roles live in input slots the vertical role has finished with
    double *const x_P1 = v_in[5], *const x_G1 = v_in[13], *const x_isf2 = v_in[14];  // netPrep, infD, infKsatV
    const int Ne = m.Ne;
    const size_t NE = (size_t)Ne;
    const size_t LD = (size_t)m.ld;  // padded leading dimension of the static [3][.] arrays
    const int lane_cell = threadIdx.x & (TILE - 1);
    const int i0 = ((int)blockIdx.x + tile0) * TILE;  // tile0: first tile of this launch (interior / boundary parts)
    const int i = i0 + lane_cell;
    const bool valid = i < Ne;
    const int ic = valid ? i : Ne - 1;  // clamped index: tail threads load something harmless
    // the river/lake kernel may be scheduled once every block of this grid is running (it fills the last wave and
    // does its state-only part; griddepcontrol.wait there holds the rest until this grid has completed)
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x >= TILE) {
        // =============================== vertical role ===============================
        // The 26 inputs of this role go global -> shared memory with per-thread 8-byte cp.async (LDGSTS): all
        // of them are in flight at once and none is parked in a register.  (With plain loads and a 64-register
        // budget the compiler spills freshly loaded values, and every spill store waits for its load, which
        // serialises the load phase into several DRAM round trips.)
        {
            const double *const src[V_NIN] = {Y + ic, Y + NE + ic, Y + 2 * NE + ic, m.satn + ic, m.eic + ic, m.netPrep + ic,
                                              m.potEvap + ic, m.potTran + ic, m.lai + ic, m.fuSurf + ic, m.fuSub + ic,
                                              m.aqd + ic, m.sy + ic, m.infD + ic, m.infKsatV + ic, m.macKsatV + ic,
                                              m.hAreaF + ic, m.thetaS + ic, m.thetaR + ic, m.thetaFC + ic, m.beta + ic,
                                              m.ksatV + ic, m.vegFrac + ic, m.impAF + ic, m.wetland + ic, m.rootReach + ic};
#pragma unroll
            for (int a = 0; a < V_NIN; a++) cp_async8(&v_in[a][lane_cell], src[a]);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        t_flv[lane_cell] = __ldg(m.flags + ic);  // issued behind the copies, not ahead of them
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#define VFL() t_flv[lane_cell]
#define VIN(a) v_in[a][lane_cell]
#define VFENCE() asm volatile("" ::: "memory")
        // The role runs in steps, each fetching only its own inputs from the staged slots and parking what a later
        // step needs back in shared memory: nothing is held in a register across the pow() calls (the 64-register
        // budget would spill it to local memory, i.e. to L2).  Order: soil first - the lateral role needs only its
        // results (P1, G1, ponding) - then the ET partition while the lateral role does the weir and the lateral
        // sums, then the three balance equations of the cell.
        if (HALO == 2 && m.h_flags && (int)blockIdx.x + tile0 >= m.n_int_tiles) {
            // a tile with ghost cells: this role reads exchanged states too - acquire the neighbours' flags (the lateral
            // role does the same for itself), then the ghost cells take (Ysurf, Yunsat, Ygw) from the halo buffer
            if (lane_cell < m.h_nflags) acquire_flag(m, lane_cell);
            bar_sync(4, TILE);
            if (VFL() & F_GHOST) {
                double gsf, gus, ggw, gkh;
                ghost_state(m, ic, VFL(), gsf, gus, ggw, gkh);
                VIN(0) = gsf; VIN(1) = gus; VIN(2) = ggw;
            }
        }
        if (VFL() & F_HEADBC) VIN(2) = m.ele_yBC[ic];
        // ---- step 1: updateElement (2 pow) ----
        SoilState st;
        if (VFL() & F_LAKE) { st.deficit = 0.; st.theta = 0.; st.satn = 1.; st.satKr = 0.; }
        else st = cell_soil_state(VIN(11), VIN(17), VIN(18), VIN(20), VIN(1), VIN(2));
        VFENCE();
        // ---- step 2: infiltration / exfiltration / recharge; hand-over ----
        {
            CellVert v;
            v.satn = 1.; v.infil = v.exfil = v.rech = 0.;
            const double ysf = VIN(0), netPrep = VIN(5);
            if (!(VFL() & F_LAKE)) {
                CellParams p;
                CellForc f;
                p.aqd = VIN(11); p.infD = VIN(13); p.infKsatV = VIN(14); p.macKsatV = VIN(15); p.hAreaF = VIN(16);
                p.thetaR = VIN(18); p.thetaFC = VIN(19); p.ksatV = VIN(21);
                f.netPrep = netPrep; f.fuSurf = VIN(9); f.fuSub = VIN(10);
                cell_soil_flux(p, f, ysf, VIN(1), VIN(2), st, v);
            }
            const double isf2 = ysf - v.infil + v.exfil;
            x_P1[lane_cell] = netPrep - v.infil + v.exfil;
            x_G1[lane_cell] = v.rech - v.exfil;
            x_isf2[lane_cell] = dmax(0., isf2);
            bar_arrive(2, 2 * TILE);  // hand-over: the lateral warps wait on barrier 2
            VIN(15) = v.infil - v.rech;  // first difference of ydot[unsat], finished after the ET step
            if (valid) {
                m.satn[i] = v.satn;
                if (DIAG) { d.qEleInfil[i] = v.infil; d.qEleExfil[i] = v.exfil; d.qEleRecharge[i] = v.rech; }
            }
        }
        VFENCE();
        // ---- step 3: ET partition (f_etFlux), with the saturation carried from the previous call ----
        CellVert v;
        v.err = 0;
        {
            const double potEvap = VIN(6);
            const unsigned fl = VFL();
            if (fl & F_LAKE) {
                v.Es = v.Eu = v.Eg = v.Tu = v.Tg = 0.; v.eic = 0.; v.iBeta = 0.;
            } else {
                CellParams p;
                CellForc f;
                p.thetaS = VIN(17); p.thetaR = VIN(18); p.vegFrac = VIN(22); p.impAF = VIN(23); p.wetland = VIN(24);
                p.rootReach = VIN(25);
                f.potEvap = potEvap; f.potTran = VIN(7); f.lai = VIN(8);
                cell_et(p, f, VIN(0), VIN(1), VIN(2), VIN(3), VIN(4), v);
            }
            if (valid) {
                m.eic[i] = v.eic;
                if (DIAG) {
                    if (fl & F_LAKE) {
                        d.qEleTrans[i] = 0.; d.qEleEvapo[i] = potEvap; d.qEleETA[i] = 0. + potEvap + 0.;
                    } else {
                        const double trans = v.Tg + v.Tu, evapo = v.Eu + v.Eg + v.Es;
                        d.qEleTrans[i] = trans; d.qEleEvapo[i] = evapo; d.qEleETA[i] = v.eic + evapo + trans;
                        d.iBeta[i] = v.iBeta;
                    }
                    d.qEs[i] = v.Es; d.qEu[i] = v.Eu; d.qEg[i] = v.Eg; d.qTu[i] = v.Tu; d.qTg[i] = v.Tg;
                }
            }
        }
        // ---- step 4: the balance equations (f_applyDY, MD_f.cpp:52-215).  The lateral role has left
        //      P1 - SurfTot/area and G1 - SubTot/area in the P1 / G1 slots (barrier 3). ----
        bar_sync(3, 2 * TILE);
        if (!valid) return;
        {
            const unsigned fl = VFL();
            const double area = t_area[lane_cell], sy = VIN(12);
            double dsf = x_P1[lane_cell] - v.Es;
            double dgw = x_G1[lane_cell] - v.Eg - v.Tg;
            if (fl & F_HEADBC) dgw = 0;
            else if (fl & F_FLUXBC) dgw += SHUD_DIVS(m.ele_QBC[i], area);
            if (fl & F_SS_SURF) dsf += SHUD_DIVS(m.qss[i], area);
            else if (fl & F_SS_GW) dgw += SHUD_DIVS(m.qss[i], area);
            dgw = SHUD_DIVS(dgw, sy);
            double dus = VIN(15) - v.Eu - v.Tu;
            dus = SHUD_DIVS(dus, sy);
            if (fl & (HALO == 2 ? (F_LAKE | F_GHOST) : F_LAKE)) { dsf = 0.; dus = 0.; dgw = 0.; }  // lake cell; ghost: its owner integrates it
            DY[i] = dsf;
            DY[NE + i] = dus;
            DY[2 * NE + i] = dgw;
            if (v.err) raise_err(m.err, v.err, i + 1);
        }
#undef VIN
#undef VFENCE
#undef VFL
        return;
    }
    // =============================== lateral role ===============================
    // Only the neighbour ids are loaded into registers (the gathers need them as addresses); every other value of
    // the cell goes global -> shared memory with cp.async, all in flight at once, and is read where it is used:
    // the role holds no parameter in a register across the three edges (no local-memory spills).
    const int nb[3] = {__ldg(m.nbr + ic), __ldg(m.nbr + LD + ic), __ldg(m.nbr + 2 * LD + ic)};
    const int q1 = __ldg(m.cell_seg_first + (i0 + TILE < Ne ? i0 + TILE : Ne));  // end of the tile's segment slots
    cp_async8(&t_sf[lane_cell], Y + ic);
    cp_async8(&t_gw[lane_cell], Y + 2 * NE + ic);
    cp_async8(&t_zs[lane_cell], m.z_surf + ic);
    cp_async8(&t_zb[lane_cell], m.z_bottom + ic);
    cp_async8(&t_dep[lane_cell], m.depression + ic);
    cp_async8(&t_fus[lane_cell], m.fuSub + ic);
    cp_async8(&t_area[lane_cell], m.area + ic);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        cp_async8(&e_B[j][lane_cell], m.edge + j * LD + ic);
        cp_async8(&e_dist[j][lane_cell], m.dist + j * LD + ic);
        cp_async8(&e_rough[j][lane_cell], m.avgRough + j * LD + ic);
    }
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&t_seg0[lane_cell])),
                 "l"(m.cell_seg_first + ic)
                 : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // effKH of this call is complete (k_effkh, PDL)
    cp_async8(&t_kh[lane_cell], m.effKH + ic);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const unsigned fl = __ldg(m.flags + ic);  // issued behind the copies, not ahead of them
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (fl & F_HEADBC) t_gw[lane_cell] = m.ele_yBC[ic];
    if (HALO && m.h_flags && (int)blockIdx.x + tile0 >= m.n_int_tiles && lane_cell < m.h_nflags) {
        // a tile that sees halo cells: one lane per neighbour partition acquires that neighbour's flag of the exchange
        // in flight before anybody in the tile reads a halo value (the barrier below orders the others behind it)
        acquire_flag(m, lane_cell);
    }
    bar_sync(1, TILE);           // the tile's own values are in shared memory (lateral warps)
    if (HALO == 2 && m.h_flags && (int)blockIdx.x + tile0 >= m.n_int_tiles) {
        if (fl & F_GHOST) {
            // ghost cell: state from the exchange, effKH evaluated here (the pre-pass saw its stale vector entry)
            double gsf, gus, ggw, gkh;
            ghost_state(m, ic, fl, gsf, gus, ggw, gkh);
            if (fl & F_HEADBC) ggw = m.ele_yBC[ic];
            t_sf[lane_cell] = gsf; t_gw[lane_cell] = ggw; t_kh[lane_cell] = gkh;
        }
        bar_sync(1, TILE);       // ... the segment pass below reads other lanes' slots
    }
    const double ysf = t_sf[lane_cell], ygw = t_gw[lane_cell], zs = t_zs[lane_cell], zb = t_zb[lane_cell];
    const double kh = t_kh[lane_cell], depression = t_dep[lane_cell], fuSub = t_fus[lane_cell];
    int err = 0;
    if (!(fl & F_LAKE)) {
        const double isf = ysf < 0. ? 0. : ysf;
        // one copy of the edge code, three trips: unrolled, the kernel outgrows the instruction cache (measured:
        // 12 % of the issue stalls were instruction fetches) - rolled it is 10 us faster
#pragma unroll 1
        for (int j = 0; j < 3; j++) {
            double qs = 0., qg = 0.;
            const int k = j == 0 ? nb[0] : (j == 1 ? nb[1] : nb[2]);
            if (k >= 0) {
                double nsf, ygw_n, zs_n, zb_n, kh_n;
                const unsigned r = (unsigned)(k - i0);
#ifndef SHUD_NBR_GLOBAL
                if (r < (unsigned)TILE && k < Ne) {  // neighbour inside the tile: shared memory (in the ragged last
                                                      // tile a halo id Ne+h also falls into [i0, i0+TILE): not a cell)
                    nsf = t_sf[r]; ygw_n = t_gw[r]; zs_n = t_zs[r]; zb_n = t_zb[r]; kh_n = t_kh[r];
                } else
#endif
                if (k < Ne) {
                    nsf = Y[k]; ygw_n = Y[2 * NE + k]; zs_n = __ldg(m.z_surf + k); zb_n = __ldg(m.z_bottom + k);
                    kh_n = m.effKH[k];
                    if (m.has_headbc && (m.flags[k] & F_HEADBC)) ygw_n = m.ele_yBC[k];
                } else if (!HALO) {
                    // Never taken: a single domain has no neighbour id >= Ne.  The branch is kept on purpose - with the
                    // test folded away at compile time ptxas schedules the gathers of the branch above differently and
                    // the kernel loses 3.5 % (measured on B200, same instruction count: 106.1 vs 102.6 us; the round-1
                    // kernel had this shape).  It mirrors the halo branch below without its arithmetic.
                    const int h = k - Ne;
                    nsf = m.h_state[2 * h]; ygw_n = m.h_state[2 * h + 1]; zs_n = m.h_zs[h]; zb_n = m.h_zb[h]; kh_n = m.h_aqd[h];
                } else {  // halo cell of a partition: state from the last halo exchange
                    halo_neighbour(m, k - Ne, nsf, ygw_n, zs_n, zb_n, kh_n);
                }
                nsf = nsf < 0. ? 0. : nsf;
                const double Bj = e_B[j][lane_cell], dj = e_dist[j][lane_cell];
                qs = edge_surface(isf, zs, nsf, zs_n, depression, dj, Bj, e_rough[j][lane_cell]);
                qg = edge_sub(ygw, zb, ygw_n, zb_n, kh, kh_n, dj, Bj);
            } else if (k <= -2) {
                const int slot = -2 - k, l = m.bank_lake[slot];
                const double yl = Y[3 * NE + m.Nr + l];
                const double nsf = yl < 0. ? 0. : yl;
                const double Bj = e_B[j][lane_cell];
                qs = weir_jtoi(m.l_zmin[l], nsf, zs, isf, zs, 0.6, Bj, 0.01);
                qg = edge_sub(ygw, zb, yl, m.l_yi0[l], kh, m.bank_kh[slot], e_dist[j][lane_cell], Bj);
            } else if (!m.close_boundary) {
                const double d2e = m.dist2edge[j * LD + ic];
                if (isf > depression) {
                    const double s = isf / d2e * 0.5;
                    if (s > 0.) qs = sqrt(s) * cbrt(isf * isf * isf * isf * isf) * e_B[j][lane_cell] / m.rough[ic];
                }
                if (ygw > depression * 10.) {
                    const double grad = ygw / d2e * 0.5;
                    if (grad > 0.) qg = kh * grad;
                }
            }
            e_B[j][lane_cell] = qs;  // the edge's statics are spent: its slots take the two fluxes
            e_dist[j][lane_cell] = qg * fuSub;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 3; j++) { e_B[j][lane_cell] = 0.; e_dist[j][lane_cell] = 0.; }
    }
    // ---- river segments of the whole tile, one lane per segment slot (dense lanes instead of a per-cell loop at
    //      ~15 % lane use): fun_Seg_sub / fun_Seg_surface, MD_RiverFlux.cpp:100-126.  The groundwater exchange needs
    //      nothing of the vertical role: it and every load of the pass run here, in the time this role would wait
    //      for the hand-over; only the weir (needs the ponding left after infiltration) comes after it. ----
    const int q0 = t_seg0[0], nsq = q1 - q0;
    const bool has_seg = lane_cell < nsq;
    int s_lc = 0, s_sgm = 0;
    double s_yr = 0., s_zr = 0., s_zbk = 0., s_cwr = 0., s_len = 0.;
    if (has_seg) {
        const int q = q0 + lane_cell;
        s_lc = __ldg(m.cs_cell + q) - i0; s_sgm = __ldg(m.cs_seg + q);
        s_yr = m.cs_yr[q];
        if (HALO == 2 && m.cs_g) { const int gs = __ldg(m.cs_g + q); if (gs >= 0) s_yr = __ldcg(halo_state(m) + m.g_roff + gs); }
        s_zr = __ldg(m.cs_zr + q); s_zbk = __ldg(m.cs_zbk + q); s_cwr = __ldg(m.cs_cwr + q);
        s_len = __ldg(m.cs_len + q);
        const double qg = flux_r2e_gw(s_yr, s_zr, t_gw[s_lc], t_zb[s_lc], t_kh[s_lc], __ldg(m.cs_ksatH + q), s_len,
                                      __ldg(m.cs_bed + q)) * t_fus[s_lc];
        m.QsegSub[s_sgm] = qg;
        sq_g[lane_cell] = qg;
    }
    bar_sync(2, 2 * TILE);  // vertical role has handed over
    if (has_seg) {
        const double qs = weir_jtoi(t_zs[s_lc], x_isf2[s_lc], s_zr, s_yr, s_zbk, s_cwr, s_len, t_dep[s_lc]);
        m.QsegSurf[s_sgm] = qs;
        sq_s[lane_cell] = qs;
    }
    for (int tq = TILE + lane_cell; tq < nsq; tq += TILE) {  // tiles with more than 128 segments (rare)
        const int q = q0 + tq;
        const int lc = __ldg(m.cs_cell + q) - i0, sgm = __ldg(m.cs_seg + q);
        double yr = m.cs_yr[q];
        if (HALO == 2 && m.cs_g) { const int gs = __ldg(m.cs_g + q); if (gs >= 0) yr = __ldcg(halo_state(m) + m.g_roff + gs); }
        const double zr = __ldg(m.cs_zr + q), len = __ldg(m.cs_len + q);
        const double qs = weir_jtoi(t_zs[lc], x_isf2[lc], zr, yr, __ldg(m.cs_zbk + q), __ldg(m.cs_cwr + q), len, t_dep[lc]);
        const double qg = flux_r2e_gw(yr, zr, t_gw[lc], t_zb[lc], t_kh[lc], __ldg(m.cs_ksatH + q), len,
                                      __ldg(m.cs_bed + q)) * t_fus[lc];
        m.QsegSurf[sgm] = qs;
        m.QsegSub[sgm] = qg;
        if (tq < SEGCAP) { sq_s[tq] = qs; sq_g[tq] = qg; }
    }
    bar_sync(1, TILE);  // segment fluxes of the tile are in shared memory
    {
        // element side of PassValue (MD_f.cpp:228-235): sum of this cell's segment fluxes, ascending segment id
        double e2rS = 0., e2rG = 0.;
        const int nseg = (int)(t_flv[lane_cell] >> NSEG_SHIFT);  // flags as the vertical role staged them (barrier 2)
        if (nseg) {
            const int seg0 = t_seg0[lane_cell], tq0 = seg0 - q0;
            for (int k = 0; k < nseg; k++) {
                const int tq = tq0 + k;
                double qs, qg;
                if (tq < SEGCAP) { qs = sq_s[tq]; qg = sq_g[tq]; }
                else { const int sgm = __ldg(m.cs_seg + seg0 + k); qs = m.QsegSurf[sgm]; qg = m.QsegSub[sgm]; }
                e2rS += -qs;
                e2rG += -qg;
            }
        }
        double surfTot = e2rS, subTot = e2rG;
        double Qs[3], Qg[3];
#pragma unroll
        for (int j = 0; j < 3; j++) { Qs[j] = e_B[j][lane_cell]; Qg[j] = e_dist[j][lane_cell]; }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            surfTot += Qs[j];
            subTot += Qg[j];
            if (not_finite(Qs[j]) || not_finite(Qg[j])) err = err > 10 ? err : 10;
        }
        // first two terms of dYsf and dYgw; the vertical role subtracts its ET terms and finishes (barrier 3)
        const double area = t_area[lane_cell];
        x_P1[lane_cell] = x_P1[lane_cell] - SHUD_DIVS(surfTot, area);
        x_G1[lane_cell] = x_G1[lane_cell] - SHUD_DIVS(subTot, area);
        bar_arrive(3, 2 * TILE);
        if (!valid) return;
        if (err) raise_err(m.err, err, i + 1);
        if (DIAG) {
#pragma unroll
            for (int j = 0; j < 3; j++) { d.QeleSurf[j * NE + i] = Qs[j]; d.QeleSub[j * NE + i] = Qg[j]; }
            d.QeleSurfTot[i] = surfTot; d.QeleSubTot[i] = subTot; d.Qe2r_Surf[i] = e2rS; d.Qe2r_Sub[i] = e2rG;
        }
    }
}

// stage of reach r as the solver sees it: the vector entry of an own reach, the exchanged stage of a ghost reach
template <int HALO>
__device__ __forceinline__ double reach_y(const DevMesh &m, const double *__restrict__ Yr, int r) {
    if (HALO == 2 && m.r_gslot) {
        const int gs = m.r_gslot[r];
        if (gs >= 0) return __ldcg(halo_state(m) + m.g_roff + gs);
    }
    return Yr[r];
}
// Manning flux of reach r towards its downstream end, everything gathered from global memory
template <int HALO>
__device__ __forceinline__ double reach_down_flux(const DevMesh &m, const double *__restrict__ Yr, int r, int *err) {
    const double yraw = reach_y<HALO>(m, Yr, r);
    const double ystg = (m.r_bc[r] > 0) ? m.r_yBC[r] : yraw;
    const int down = m.r_down[r];
    double y_dn = 0., depth_dn = 0., slope_dn = 0.;
    if (down >= 0) {
        y_dn = (m.r_bc[down] > 0) ? m.r_yBC[down] : reach_y<HALO>(m, Yr, down);
        depth_dn = m.r_depth[down];
        slope_dn = m.r_slope[down];
    }
    // r_down device coding: >=0 downstream reach (device id); <0 the reference's outlet code
    return river_down(yraw, ystg, m.r_w0[r], m.r_bank[r], m.r_len[r], m.r_slope[r], m.r_depth[r], m.r_rough[r],
                      m.r_dist[r], down >= 0 ? 1 : down, m.r_toLake[r], y_dn, depth_dn, slope_dn, err);
}

// fixed-shape block sum (deterministic): warp shuffle tree, then warp 0 over the warp partials
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sm) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < NT / 32) ? sm[threadIdx.x] : 0.;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// ---------------------------------------------------------------------------------------------
// K2: blocks [0, nb_riv) - one thread per reach: routing, river side of PassValue, stage equation
//     (Flux_RiverDown MD_RiverFlux.cpp:5-63, PassValue MD_f.cpp:228-240, f_applyDY MD_f.cpp:157-179);
//     blocks [nb_riv, nb_riv+Nl) - one block per lake (MD_f.cpp:16-17,44-47,180-191).
// ---------------------------------------------------------------------------------------------
template <bool DIAG, int HALO>
__device__ __forceinline__ void river_lake_body(const DevMesh &m, const DevDiag &d, const double *__restrict__ Y,
                                                double *__restrict__ DY, int nb_riv, double *sm) {
    const size_t NE = (size_t)m.Ne;
    const size_t LD = (size_t)m.ld;
    const double *Yr = Y + 3 * NE;
    if (HALO == 2 && m.h_flags && m.r_gslot) {
        // ghost reaches: this kernel reads exchanged stages too; one lane per neighbour acquires its flag (long set)
        if ((int)threadIdx.x < m.h_nflags) {
            const unsigned long long e = *m.h_epoch + 1ull;
            unsigned long long t0;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
            while (ld_acquire_sys(m.h_flags + threadIdx.x) < e) {
                __nanosleep(200);
                unsigned long long t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > 5000000000ull) { raise_err(m.err, SHUD_ERR_P2P_TIMEOUT, (int)threadIdx.x + 1); break; }
            }
        }
        __syncthreads();
    }
    if ((int)blockIdx.x < nb_riv) {
        const int r = blockIdx.x * blockDim.x + threadIdx.x;
        if (r >= m.Nr || (HALO == 2 && m.r_gslot && m.r_gslot[r] >= 0)) {
            // past the end, or a ghost reach: its owner integrates it, its entry of ydot is 0 here
            if (r < m.Nr) DY[3 * NE + r] = 0.;
            return;
        }
        const int s0 = m.r_seg_ptr[r], s1 = m.r_seg_ptr[r + 1], u0 = m.r_up_ptr[r], u1 = m.r_up_ptr[r + 1];
        const int bc = m.r_bc[r];
        const double yraw = Yr[r], w0 = m.r_w0[r], bank = m.r_bank[r], len = m.r_len[r];
        const double qbc = (bc < 0) ? m.r_qBC[r] : 0.;
        const RivGeom g = riv_geom(yraw, w0, bank);
        // Flux_RiverDown of this reach and of its upstream reaches (re-evaluated rather than exchanged: state only)
        int err = 0;
        const double qdown = reach_down_flux<HALO>(m, Yr, r, &err);
        double up = 0.;
        for (int k = u0; k < u1; k++) up += -reach_down_flux<HALO>(m, Yr, m.r_up_idx[k], &err);
        if (err) raise_err(m.err, err, r + 1);
        // everything above needs statics and the state only; the segment fluxes come from the cell kernel
        // (programmatic dependent launch: this grid starts in the cell kernel's last wave and waits here)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        double surf = 0., sub = 0.;
        for (int s = s0; s < s1; s++) {
            surf += m.QsegSurf[s];
            sub += m.QsegSub[s];
        }
        double dy;
        if (bc > 0) {
            dy = 0.;
        } else {
            dy = (-up - surf - sub - qdown + qbc) / len;
            if (dy < -1. * g.csArea) dy = -1. * g.csArea;
            dy = dA_to_dY(dy, g.topWidth, bank);
        }
        DY[3 * NE + r] = dy;
        if (DIAG) { d.QrivSurf[r] = surf; d.QrivSub[r] = sub; d.QrivUp[r] = up; d.QrivDown[r] = qdown; }
        return;
    }
    // ---- lake l ----
    const int l = blockIdx.x - nb_riv;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const double yl = Y[3 * NE + m.Nr + l];
    double qs = 0., qg = 0., qin = 0.;
    for (int k = m.l_bank_ptr[l] + threadIdx.x; k < m.l_bank_ptr[l + 1]; k += blockDim.x) {
        const int i = m.bank_cell[k], j = m.bank_j[k];
        const unsigned fl = m.flags[i];
        double ysf = Y[i];
        const double isf = ysf < 0. ? 0. : ysf, zs = m.z_surf[i];
        const double ygw = (fl & F_HEADBC) ? m.ele_yBC[i] : Y[2 * NE + i];
        const double B = m.edge[j * LD + i];
        qs += weir_jtoi(m.l_zmin[l], yl < 0. ? 0. : yl, zs, isf, zs, 0.6, B, 0.01);
        // QLakeSub takes Q before the fu_Sub factor (MD_ElementFlux.cpp:121 precedes :153)
        qg += edge_sub(ygw, m.z_bottom[i], yl, m.l_yi0[l], m.effKH[i], m.bank_kh[k], m.dist[j * LD + i], B);
    }
    for (int k = m.l_rin_ptr[l] + threadIdx.x; k < m.l_rin_ptr[l + 1]; k += blockDim.x)
    { int e2 = 0; qin += reach_down_flux<HALO>(m, Yr, m.l_rin_idx[k], &e2); }
    qs = block_sum<128>(qs, sm);
    qg = block_sum<128>(qg, sm);
    qin = block_sum<128>(qin, sm);
    if (threadIdx.x == 0) {
        const int b0 = m.l_bptr[l], b1 = m.l_bptr[l + 1];
        const double area = lake_toparea(m.l_by + b0, m.l_ba + b0, b1 - b0, yl + m.l_zmin[l]);
        const double prcp = m.l_prcp[l];
        double evap = dmin(m.l_evap_raw[l], prcp + yl);
        evap = dmax(0., evap);
        const double qout = 0.;  // QLakeRivOut is never fed in the reference (MD_update.cpp:184)
        DY[3 * NE + m.Nr + l] = prcp - evap + (qin - qout + qg + qs) / area;
        if (DIAG) {
            d.y2LakeArea[l] = area; d.QLakeSurf[l] = qs; d.QLakeSub[l] = qg; d.QLakeRivIn[l] = qin;
            d.QLakeRivOut[l] = qout; d.qLakeEvap[l] = evap; d.qLakePrcp[l] = prcp;
        }
    }
}
template <bool DIAG, int HALO>
__global__ void __launch_bounds__(128) k_river_lake(DevMesh m, DevDiag d, const double *__restrict__ Y,
                                                    double *__restrict__ DY, int nb_riv) {
    __shared__ double sm[8];
    river_lake_body<DIAG, HALO>(m, d, Y, DY, nb_riv, sm);
    if (HALO == 1 && m.h_flags && blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) {
        // no ghost reach: no block of this grid reads the epoch word, the last thread advances it once the cell kernel
        // (whose halo tiles read it) has completed
        asm volatile("griddepcontrol.wait;" ::: "memory");
        *m.h_epoch = *m.h_epoch + 1ull;
    }
    if (HALO == 2 && m.h_flags) {
        // peer-to-peer exchange: the epoch word advances when the LAST block of this grid is through - every block reads
        // it (parity of the halo buffer) - and the cell kernel, whose halo tiles read it too, has completed
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            __threadfence();
            if (atomicAdd(m.h_done, 1u) == gridDim.x - 1) {
                *m.h_done = 0u;
                *m.h_epoch = *m.h_epoch + 1ull;
            }
        }
    }
}

// send side of the halo exchange: (Ysurf, Ygw) of the listed owned cells
__global__ void k_pack_halo(const double *__restrict__ Y, const int *__restrict__ idx, int n, int Ne, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    out[2 * k] = Y[i];
    out[2 * k + 1] = Y[2 * (size_t)Ne + i];
}

// Print_Ctrl::PrintData on the device: acc += value (Model_Control.cpp:933-935)
__global__ void k_accumulate(double *__restrict__ acc, const double *__restrict__ v, size_t n) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        acc[k] += v[k];
}
__global__ void k_scale_copy(double *__restrict__ dst, double *__restrict__ acc, double s, size_t n) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        dst[k] = acc[k] * s;  // buffer *= tau / NumUpdate (Model_Control.cpp:944-946); the reset follows the download
    }
}

// carried state from y (Model_Data::updateforcing -> updateElement for every cell, MD_ET.cpp:14-19)
__global__ void k_prime(DevMesh m, const double *__restrict__ Y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.Ne) return;
    // updateforcing reads uYgw, which f_update set to the BC head for iBC > 0 cells (MD_update.cpp:114-118)
    const double ygw = (m.flags[i] & F_HEADBC) ? m.ele_yBC[i] : Y[2 * (size_t)m.Ne + i];
    m.satn[i] = cell_satn(m.aqd[i], m.thetaS[i], m.thetaR[i], Y[(size_t)m.Ne + i], ygw);
}

// Model_Data::summary (MD_update.cpp:190-216): the state the host prints / checkpoints is the solver vector with the
// groundwater head of head-BC cells and the stage of stage-BC reaches replaced by their BC values (device order)
__global__ void k_summary(DevMesh m, const double *__restrict__ Y, double *__restrict__ out, size_t NY) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= NY) return;
    const size_t NE = (size_t)m.Ne;
    double v = Y[k];
    if (k >= 2 * NE && k < 3 * NE) {
        if (m.flags[k - 2 * NE] & F_HEADBC) v = m.ele_yBC[k - 2 * NE];
    } else if (k >= 3 * NE && k < 3 * NE + (size_t)m.Nr) {
        if (m.r_bc[k - 3 * NE] > 0) v = m.r_yBC[k - 3 * NE];
    }
    out[k] = v;
}

// reference order <-> device order (cells x3 blocks, reaches; lakes keep their order)
__global__ void k_to_dev(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ cperm,
                         const int *__restrict__ rperm, int Ne, int Nr, int Nl) {
    const size_t n = 3 * (size_t)Ne + Nr + Nl;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        size_t s;
        if (k < 3 * (size_t)Ne) {
            const size_t b = k / Ne, i = k - b * Ne;
            s = b * Ne + cperm[i];
        } else if (k < 3 * (size_t)Ne + Nr) {
            s = 3 * (size_t)Ne + rperm[k - 3 * (size_t)Ne];
        } else {
            s = k;
        }
        dst[k] = src[s];
    }
}
__global__ void k_from_dev(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ cperm,
                           const int *__restrict__ rperm, int Ne, int Nr, int Nl) {
    const size_t n = 3 * (size_t)Ne + Nr + Nl;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        size_t s;
        if (k < 3 * (size_t)Ne) {
            const size_t b = k / Ne, i = k - b * Ne;
            s = b * Ne + cperm[i];
        } else if (k < 3 * (size_t)Ne + Nr) {
            s = 3 * (size_t)Ne + rperm[k - 3 * (size_t)Ne];
        } else {
            s = k;
        }
        dst[s] = src[k];
    }
}

// Hilbert index of (x,y) on a 2^16 x 2^16 grid
uint64_t hilbert_d(uint32_t x, uint32_t y) {
    uint64_t dd = 0;
    for (uint32_t s = 1u << 15; s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
        dd += (uint64_t)s * s * ((3 * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = 65535u - x; y = 65535u - y; }
            std::swap(x, y);
        }
    }
    return dd;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side: context
// ---------------------------------------------------------------------------------------------
struct shud_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int Ne = 0, Nr = 0, Ns = 0, Nl = 0, Nhalo = 0;
    int ld = 0;  // Ne rounded up to a multiple of 128
    int64_t NY = 0;
    DevMesh m{};
    DevDiag diag{};
    DevLand land{};
    bool has_land = false;
    // pinned (mapped) staging of the per-step tables, two halves used alternately: the host fills one while the
    // device may still read the other (land_ev[k]: the last k_land_tables that read half k)
    double *land_stage = nullptr;
    double *land_stage_dev = nullptr;  // its device alias
    size_t land_ntab = 0;
    cudaEvent_t land_ev[2] = {nullptr, nullptr};
    int land_ev_set[2] = {0, 0};
    int land_flip = 0;
    double cryo_tstart = -9999., cryo_nday = 0.;  // _AccTemp::Time_start, N_of_day (identical for every cell)
    int cryo_size_s = 0, cryo_head_s = 0, cryo_size_b = 0, cryo_head_b = 0;
    bool diag_alloc = false;
    DevDiag acc{};          // output accumulators (same shapes as diag) + effKH/satn/Qseg copies
    double *acc_effKH = nullptr, *acc_satn = nullptr, *acc_QsegSurf = nullptr, *acc_QsegSub = nullptr;
    double *acc_tmp = nullptr;
    bool acc_alloc = false;
    int num_update = 0;
    std::vector<void *> allocs;
    char *arena = nullptr;          // one block for the lateral role's statics (optional persisting-L2 window)
    size_t arena_off = 0, arena_cap = 0;
    bool arena_on = false;
    std::vector<int> cperm, rperm, sperm;  // device id -> reference id (0-based)
    std::vector<int> cinv, rinv;           // reference id -> device id
    int *d_cperm = nullptr, *d_rperm = nullptr;
    double *y_stage = nullptr, *y_dev = nullptr, *ydot_dev = nullptr;  // for the host-pointer entry point
    std::vector<int> lake_cells;  // reference ids of lake cells, ascending
    std::vector<int> lake_of_cell;
    std::vector<int> lake_nele;
    double *h_pinned = nullptr;  // staging for forcing uploads (pinned)
    int *h_err = nullptr;        // error word (mapped pinned; m.err is its device alias)
    double *f_stage = nullptr;   // device staging of a forcing upload: 9 cell columns + 2 reach columns, reference order
    std::vector<double *> flush_tmp;  // persistent scratch of shud_b200_output_flush
    unsigned long graph_clock = 0;    // LRU stamp of the graph cache
    size_t h_pinned_n = 0;
    bool has_ebc_arrays = false;
    // partition: tiles whose cells see no halo cell (interior) / the others (boundary) - overlap of the exchange
    int n_int_tiles = 0, n_bnd_tiles = 0;
    cudaEvent_t ev_kh = nullptr, ev_bnd = nullptr;  // effKH of the owned cells done / boundary tiles done (rhs_boundary_dev)
    // halo exchange over NCCL driven from here (shud_b200_comm_init / shud_b200_exchange_plan / shud_b200_rhs_exchange_dev)
    void *nccl_dl = nullptr, *nccl_comm = nullptr;
    int (*nccl_send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*nccl_recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*nccl_group_start)() = nullptr, (*nccl_group_end)() = nullptr;
    int (*nccl_comm_destroy)(void *) = nullptr;
    int (*nccl_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    double *ar_dev = nullptr;         // scratch of the scalar allreduce (SHUD_AR_MAX doubles)
    cudaStream_t xstream = nullptr;   // the exchange (and the boundary tiles) run here, beside the interior tiles
    cudaEvent_t ev_pack = nullptr;
    std::vector<int> x_peer, x_scount, x_rcount;  // per neighbour partition: rank, cells sent, halo cells received
    int *x_sidx = nullptr;            // device-order ids of the cells sent, concatenated by peer
    double *x_sbuf = nullptr, *x_hstate = nullptr;  // packed (Ysurf, Ygw) pairs out / halo state in
    int x_nsend = 0, x_cap_send = 0;
    int n_ghost_cells = 0, n_ghost_reaches = 0;
    int force_halo = 0;               // SHUD_FORCE_HALO=1: a single domain runs the partition kernels (developer A/B)
    // the exchange as doubles grouped by (neighbour, kind) - kinds: 0 halo pairs, 1 ghost-cell triples, 2 ghost-reach stages
    std::vector<int> x_scount3, x_rcount3;   // [npeers][3]
    int *x_items = nullptr;                  // flat device-order indices into the state vector of what I send
    int x_nitems = 0;
    // peer-to-peer exchange (shud_b200_p2p_export / _connect)
    int ar_nranks = 0, ar_rank = 0;       // mailboxes of the vector reductions' in-kernel allreduce, one per rank
    void *ar_box[16] = {nullptr};
    void *p2p_block = nullptr;            // [flags: P2P_MAXPEER u64 | epoch | count | pad to 256 B][buffer 0][buffer 1]
    size_t p2p_stride = 0;                // bytes of one halo buffer (multiple of 256)
    std::vector<void *> p2p_opened;       // neighbours' blocks mapped through CUDA IPC
    P2PTable p2p{};
    int use_p2p = 0;
    int use_xgraph = 1;               // SHUD_XGRAPH: rhs_exchange_dev replayed as one CUDA graph per (y, ydot)
    // CUDA graphs of the solver-mode launch sequence, one per (y, ydot) pointer pair CVODE hands in
    int use_graph = 1;
    int use_pdl = 1;   // programmatic dependent launch of the cell kernel behind k_effkh (SHUD_PDL)
    struct GraphEntry { const double *y; double *yd; cudaGraphExec_t exec; unsigned long used; };
    std::vector<GraphEntry> graphs, xgraphs;
    // graphs of shud_b200_rhs_dq_dev: keyed by every pointer of the call and sigma
    struct DqGraph { DqArgs a; double *yd; cudaGraphExec_t exec; unsigned long used; };
    std::vector<DqGraph> dqgraphs;
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            fprintf(stderr, "[shud_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SHUD_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

namespace {

template <class T>
T *dev_alloc(shud_ctx *c, size_t n) {
    void *p = nullptr;
    if (n == 0) n = 1;
    if (c->arena_on) {
        const size_t b = (n * sizeof(T) + 255) & ~(size_t)255;
        if (c->arena_off + b <= c->arena_cap) {
            p = c->arena + c->arena_off;
            c->arena_off += b;
            return (T *)p;
        }
    }
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return nullptr;
    c->allocs.push_back(p);
    return (T *)p;
}
template <class T>
T *dev_upload(shud_ctx *c, const std::vector<T> &h) {
    T *p = dev_alloc<T>(c, h.size());
    if (p && !h.empty()) cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}
// gather a per-cell host array into device order and upload
// static divisors are stored as reciprocals under SHUD_RCP (see shud_phys.cuh)
#ifdef SHUD_RCP
inline double divisor(double x) { return 1.0 / x; }
#else
inline double divisor(double x) { return x; }
#endif
// per-cell arrays are padded to c->ld (multiple of 128) so that a 128-cell tile is always one whole,
// 16-byte aligned 1 KB slice (TMA bulk copies); the pad holds 1.0 (harmless in every formula)
const double *up_cell(shud_ctx *c, const double *src, bool as_divisor = false) {
    std::vector<double> h(c->ld, 1.0);
    for (int i = 0; i < c->Ne; i++) h[i] = as_divisor ? divisor(src[c->cperm[i]]) : src[c->cperm[i]];
    return dev_upload(c, h);
}
const double *up_edge(shud_ctx *c, const double *src, bool as_divisor = false) {
    std::vector<double> h(3 * (size_t)c->ld, 1.0);
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < c->Ne; i++) {
            const double x = src[(size_t)j * c->Ne + c->cperm[i]];
            h[(size_t)j * c->ld + i] = as_divisor ? divisor(x) : x;
        }
    return dev_upload(c, h);
}
const double *up_riv(shud_ctx *c, const double *src) {
    std::vector<double> h(c->Nr);
    for (int i = 0; i < c->Nr; i++) h[i] = src[c->rperm[i]];
    return dev_upload(c, h);
}

}  // namespace

extern "C" {

static void drop_graphs(shud_ctx *c);
// graph cache of 32 (y, ydot) pointer pairs: the least recently launched one makes room (CVODE alternates between a
// handful of work vectors; a long-lived cache entry must not be lost because a 33rd pair showed up once)
static void evict_lru(shud_ctx *c, std::vector<shud_ctx::GraphEntry> &cache) {
    (void)c;
    if (cache.size() < 32) return;
    size_t lru = 0;
    for (size_t k = 1; k < cache.size(); k++)
        if (cache[k].used < cache[lru].used) lru = k;
    cudaGraphExecDestroy(cache[lru].exec);
    cache.erase(cache.begin() + lru);
}

int shud_b200_create(const shud_mesh *M, int device, shud_ctx **out) {
    return shud_b200_create_partition(M, nullptr, device, out);
}

int shud_b200_create_partition(const shud_mesh *M, const shud_halo *H, int device, shud_ctx **out) {
    if (!M || !out || M->Ne <= 0) return SHUD_ERR_ARG;
    const int Nhalo = (H && H->Nhalo > 0) ? H->Nhalo : 0;
    const int ngc = H ? std::max(0, (int)H->n_ghost_cells) : 0, ngr = H ? std::max(0, (int)H->n_ghost_reaches) : 0;
    if (ngc > M->Ne || ngr > M->Nr) return SHUD_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) return SHUD_ERR_NO_DEVICE;
    CK(cudaSetDevice(device));
    shud_ctx *c = new shud_ctx();
    c->device = device;
    const int Ne = c->Ne = M->Ne, Nr = c->Nr = M->Nr, Ns = c->Ns = M->Ns, Nl = c->Nl = M->Nl;
    c->NY = 3 * (int64_t)Ne + Nr + Nl;
    c->ld = ((Ne + 127) / 128) * 128;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));

    // ---- cell order: Hilbert curve over centroids (if given) ----
    c->cperm.resize(Ne);
    std::iota(c->cperm.begin(), c->cperm.end(), 0);
    if (M->x && M->y && Ne > 1) {
        double x0 = M->x[0], x1 = x0, y0 = M->y[0], y1 = y0;
        for (int i = 1; i < Ne; i++) {
            x0 = std::min(x0, M->x[i]); x1 = std::max(x1, M->x[i]);
            y0 = std::min(y0, M->y[i]); y1 = std::max(y1, M->y[i]);
        }
        const double span = std::max(std::max(x1 - x0, y1 - y0), 1e-30);
        std::vector<uint64_t> key(Ne);
        for (int i = 0; i < Ne; i++) {
            const uint32_t hx = (uint32_t)std::min(65535.0, (M->x[i] - x0) / span * 65535.0);
            const uint32_t hy = (uint32_t)std::min(65535.0, (M->y[i] - y0) / span * 65535.0);
            key[i] = hilbert_d(hx, hy);
        }
        std::stable_sort(c->cperm.begin(), c->cperm.end(), [&](int a, int b) { return key[a] < key[b]; });
    }
    // ---- a partition keeps the tiles that see a halo cell behind the others: the device order is
    //      [interior tiles | boundary tiles (+ the ragged last tile)], so each part of the RHS is one contiguous
    //      tile range (shud_b200_rhs_interior_dev / _boundary_dev) ----
    {
        const int ntile = (Ne + TILE - 1) / TILE, nfull = Ne / TILE;
        c->n_int_tiles = ntile; c->n_bnd_tiles = 0;
        if (Nhalo > 0 || ngc > 0 || ngr > 0) {
            // tiles that read exchanged data: they see a halo cell, hold a ghost cell, or have a segment on a ghost reach
            std::vector<char> on_ghost_reach(Ne, 0);
            for (int sg = 0; sg < Ns; sg++)
                if (M->seg_iRiv[sg] - 1 >= Nr - ngr && M->seg_iEle[sg] >= 1 && M->seg_iEle[sg] <= Ne) on_ghost_reach[M->seg_iEle[sg] - 1] = 1;
            std::vector<char> isb(ntile, 0);
            for (int i = 0; i < Ne; i++) {
                const int o = c->cperm[i];
                if (o >= Ne - ngc || on_ghost_reach[o]) isb[i / TILE] = 1;
                for (int j = 0; j < 3; j++)
                    if (M->nabr[(size_t)j * Ne + o] - 1 >= Ne) isb[i / TILE] = 1;
            }
            std::vector<int> order;
            order.reserve(Ne);
            int n_int = 0;
            for (int pass = 0; pass < 2; pass++)
                for (int t = 0; t < nfull; t++)
                    if ((int)isb[t] == pass) {
                        if (!pass) n_int++;
                        for (int k = 0; k < TILE; k++) order.push_back(c->cperm[(size_t)t * TILE + k]);
                    }
            for (int i = nfull * TILE; i < Ne; i++) order.push_back(c->cperm[i]);
            c->cperm.swap(order);
            c->n_int_tiles = n_int; c->n_bnd_tiles = ntile - n_int;
        }
    }
    c->cinv.resize(Ne);
    for (int i = 0; i < Ne; i++) c->cinv[c->cperm[i]] = i;

    // ---- reach order: follow the cells their segments touch ----
    c->rperm.resize(Nr);
    std::iota(c->rperm.begin(), c->rperm.end(), 0);
    {
        std::vector<int64_t> key(Nr, INT64_MAX / 2);
        for (int s = 0; s < Ns; s++) {
            const int r = M->seg_iRiv[s] - 1, e = M->seg_iEle[s] - 1;
            if (r < 0 || r >= Nr || e < 0 || e >= Ne) { delete c; return SHUD_ERR_ARG; }
            key[r] = std::min<int64_t>(key[r], c->cinv[e]);
        }
        std::stable_sort(c->rperm.begin(), c->rperm.end(), [&](int a, int b) { return key[a] < key[b]; });
    }
    c->rinv.resize(Nr);
    for (int i = 0; i < Nr; i++) c->rinv[c->rperm[i]] = i;

    // ---- segment order: grouped by reach (device id), ascending reference id inside ----
    c->sperm.resize(Ns);
    std::iota(c->sperm.begin(), c->sperm.end(), 0);
    std::stable_sort(c->sperm.begin(), c->sperm.end(),
                     [&](int a, int b) { return c->rinv[M->seg_iRiv[a] - 1] < c->rinv[M->seg_iRiv[b] - 1]; });
    std::vector<int> sinv(Ns);
    for (int s = 0; s < Ns; s++) sinv[c->sperm[s]] = s;

    DevMesh &m = c->m;
    m.Ne = Ne; m.Nr = Nr; m.Ns = Ns; m.Nl = Nl;
    m.ld = c->ld;
    const int LDh = c->ld;
    m.close_boundary = M->close_boundary;
    {
        if (getenv("SHUD_PDL")) c->use_pdl = atoi(getenv("SHUD_PDL"));
        if (getenv("SHUD_FORCE_HALO")) c->force_halo = atoi(getenv("SHUD_FORCE_HALO"));
        if (getenv("SHUD_GRAPH")) c->use_graph = atoi(getenv("SHUD_GRAPH"));
        if (getenv("SHUD_XGRAPH")) c->use_xgraph = atoi(getenv("SHUD_XGRAPH"));
    }

    // ---- static per-cell arrays ----
    {
        // the statics the lateral role streams every call, contiguous: SHUD_L2_PERSIST=1 pins them in L2
        c->arena_cap = (size_t)12 * ((size_t)c->ld * sizeof(double) + 256);
        void *a = nullptr;
        if (cudaMalloc(&a, c->arena_cap) == cudaSuccess) { c->arena = (char *)a; c->allocs.push_back(a); }
        else { cudaGetLastError(); c->arena_cap = 0; }
    }
    c->arena_on = c->arena != nullptr;
    m.area = up_cell(c, M->area, true); m.z_surf = up_cell(c, M->z_surf); m.z_bottom = up_cell(c, M->z_bottom);
    c->arena_on = false;
    m.depression = up_cell(c, M->depression); m.aqd = up_cell(c, M->AquiferDepth); m.sy = up_cell(c, M->Sy, true);
    m.infD = up_cell(c, M->infD); m.infKsatV = up_cell(c, M->infKsatV); m.macKsatV = up_cell(c, M->macKsatV);
    m.hAreaF = up_cell(c, M->hAreaF); m.thetaS = up_cell(c, M->ThetaS); m.thetaR = up_cell(c, M->ThetaR);
    m.thetaFC = up_cell(c, M->ThetaFC); m.beta = up_cell(c, M->Beta); m.ksatH = up_cell(c, M->KsatH);
    m.ksatV = up_cell(c, M->KsatV); m.macKsatH = up_cell(c, M->macKsatH); m.macD = up_cell(c, M->macD);
    m.vAreaF = up_cell(c, M->geo_vAreaF); m.vegFrac = up_cell(c, M->VegFrac); m.impAF = up_cell(c, M->ImpAF);
    m.wetland = up_cell(c, M->WetlandLevel); m.rootReach = up_cell(c, M->RootReachLevel);
    m.rough = up_cell(c, M->Rough); m.qss = up_cell(c, M->QSS);
    c->arena_on = c->arena != nullptr;
    m.edge = up_edge(c, M->edge); m.dist = up_edge(c, M->Dist2Nabor, true);
    m.avgRough = up_edge(c, M->avgRough, true);
    c->arena_on = false;
    m.dist2edge = up_edge(c, M->Dist2Edge);
    if (c->arena && getenv("SHUD_L2_PERSIST") && atoi(getenv("SHUD_L2_PERSIST")) > 0) {
        int max_persist = 0, max_win = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, device);
        const size_t setaside = std::min((size_t)max_persist, c->arena_off);
        const size_t win = std::min((size_t)max_win, c->arena_off);
        if (setaside > 0 && win > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setaside) == cudaSuccess) {
            cudaStreamAttrValue av = {};
            av.accessPolicyWindow.base_ptr = c->arena;
            av.accessPolicyWindow.num_bytes = win;
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setaside / (double)win);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaError_t e = cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            fprintf(stderr, "[shud_b200] L2 persisting window: %zu MB of %zu MB, set-aside %zu MB (max %d MB): %s\n",
                    win >> 20, c->arena_off >> 20, setaside >> 20, max_persist >> 20, cudaGetErrorString(e));
        }
        cudaGetLastError();
    }

    // ---- topology, flags, bank edges ----
    const bool lakeon = M->lakeon != 0 && Nl > 0;
    std::vector<unsigned> flags(LDh, 0u);
    std::vector<int> nbr(3 * (size_t)LDh, -1);
    std::vector<int> bank_cell, bank_j, bank_lake;
    std::vector<double> bank_kh;
    // bank edges in ascending reference (cell, edge) order, grouped by lake afterwards
    struct Bank { int lake, ref_cell, j; double kh; };
    std::vector<Bank> banks;
    int has_headbc = 0;
    c->lake_of_cell.assign(Ne, -1);
    for (int o = 0; o < Ne; o++) {  // o = reference id
        const int i = c->cinv[o];
        unsigned f = 0;
        const bool is_lake = lakeon && M->iLake[o] > 0;
        if (M->iLake[o] > 0) f |= F_LAKE;  // f_applyDY tests iLake alone (MD_f.cpp:146)
        if (is_lake) { c->lake_cells.push_back(o); c->lake_of_cell[o] = M->iLake[o] - 1; }
        if (M->iBC[o] > 0) { f |= F_HEADBC; has_headbc = 1; }
        else if (M->iBC[o] < 0) f |= F_FLUXBC;
        if (M->iSS[o] > 0) f |= F_SS_SURF;
        else if (M->iSS[o] < 0) f |= F_SS_GW;
        if (o >= Ne - ngc) f |= F_GHOST;
        flags[i] = f;
        for (int j = 0; j < 3; j++) {
            const int nb = M->nabr[(size_t)j * Ne + o] - 1;
            const int lk = M->lakenabr ? M->lakenabr[(size_t)j * Ne + o] - 1 : -1;
            int v = -1;
            if (lk >= 0 && lk < Nl && !is_lake) {
                // effKH of the lake-side cell is its KsatH (updateLakeElement)
                banks.push_back({lk, o, j, (nb >= 0 && nb < Ne) ? M->KsatH[nb] : 0.0});
                v = -2;  // slot patched below
            } else if (nb >= 0 && nb < Ne) {
                v = c->cinv[nb];
            } else if (nb >= Ne && nb < Ne + Nhalo) {
                v = nb;  // halo cells keep their place after the owned cells
            }
            nbr[(size_t)j * LDh + i] = v;
        }
    }
    if (M->iLake && !lakeon)
        for (int o = 0; o < Ne; o++)
            if (M->iLake[o] > 0) { /* lake module off but cell tagged: treated as land in f_loop */ }
    // group banks by lake (stable: keeps ascending (cell, edge) inside a lake)
    std::stable_sort(banks.begin(), banks.end(), [](const Bank &a, const Bank &b) { return a.lake < b.lake; });
    std::vector<int> l_bank_ptr(Nl + 1, 0);
    for (size_t k = 0; k < banks.size(); k++) {
        const Bank &b = banks[k];
        const int i = c->cinv[b.ref_cell];
        nbr[(size_t)b.j * LDh + i] = -2 - (int)k;
        bank_cell.push_back(i); bank_j.push_back(b.j); bank_lake.push_back(b.lake); bank_kh.push_back(b.kh);
        l_bank_ptr[b.lake + 1]++;
    }
    for (int l = 0; l < Nl; l++) l_bank_ptr[l + 1] += l_bank_ptr[l];
    m.nbank = (int)banks.size();
    m.has_headbc = has_headbc;

    // ---- cell -> segments (ascending reference segment id) ----
    std::vector<int> nseg(Ne, 0), cell_seg_first(LDh + 1, 0), cell_seg_idx(Ns), cs_cell_h(Ns);
    for (int s = 0; s < Ns; s++) nseg[c->cinv[M->seg_iEle[s] - 1]]++;
    {
        int acc = 0;
        for (int i = 0; i < Ne; i++) {
            cell_seg_first[i] = acc;
            for (int k = 0; k < nseg[i]; k++) cs_cell_h[acc + k] = i;
            acc += nseg[i];
        }
        for (int i = Ne; i <= LDh; i++) cell_seg_first[i] = acc;
        std::vector<int> fill(Ne, 0);
        for (int s = 0; s < Ns; s++) {  // ascending reference id
            const int i = c->cinv[M->seg_iEle[s] - 1];
            cell_seg_idx[cell_seg_first[i] + fill[i]++] = sinv[s];
        }
        for (int i = 0; i < Ne; i++) {
            if (nseg[i] > 0xFFFFFF) { delete c; return SHUD_ERR_ARG; }
            flags[i] |= (unsigned)nseg[i] << NSEG_SHIFT;
        }
    }
    m.flags = dev_upload(c, flags); m.nbr = dev_upload(c, nbr);
    m.cell_seg_first = dev_upload(c, cell_seg_first);
    {
        std::vector<int> cs_seg(Ns), cs_riv(Ns), cs_bc(Ns);
        std::vector<double> cs_len(Ns), cs_cwr(Ns), cs_depth(Ns), cs_zbank(Ns), cs_ksatH(Ns), cs_bed(Ns);
        std::vector<double> cs_zr(Ns), cs_zbk(Ns);
        for (int q = 0; q < Ns; q++) {
            const int sd = cell_seg_idx[q];   // segment, device order
            const int so = c->sperm[sd];      // segment, reference id
            const int ro = M->seg_iRiv[so] - 1;
            cs_seg[q] = sd; cs_riv[q] = c->rinv[ro]; cs_bc[q] = M->riv_BC[ro];
            cs_len[q] = M->seg_length[so]; cs_cwr[q] = M->seg_Cwr[so]; cs_depth[q] = M->riv_depth[ro];
            cs_zbank[q] = M->riv_zbank[ro]; cs_ksatH[q] = M->riv_KsatH[ro]; cs_bed[q] = M->riv_BedThick[ro];
            // the two sums fun_Seg_surface forms from statics (MD_RiverFlux.cpp:104-109), same IEEE operations
            const double zs_cell = M->z_surf[c->cperm[cs_cell_h[q]]];
            cs_zr[q] = zs_cell - cs_depth[q];
            cs_zbk[q] = zs_cell + cs_zbank[q];
        }
        m.cs_seg = dev_upload(c, cs_seg); m.cs_riv = dev_upload(c, cs_riv); m.cs_bc = dev_upload(c, cs_bc);
        m.cs_cell = dev_upload(c, cs_cell_h);
        m.cs_len = dev_upload(c, cs_len); m.cs_cwr = dev_upload(c, cs_cwr); m.cs_depth = dev_upload(c, cs_depth);
        m.cs_zbank = dev_upload(c, cs_zbank); m.cs_ksatH = dev_upload(c, cs_ksatH); m.cs_bed = dev_upload(c, cs_bed);
        m.cs_zr = dev_upload(c, cs_zr); m.cs_zbk = dev_upload(c, cs_zbk);
        m.cs_yr = dev_alloc<double>(c, Ns);
        m.cs_g = nullptr; m.g_cslot = nullptr; m.r_gslot = nullptr;
        m.g_coff = 2 * Nhalo; m.g_roff = 2 * Nhalo + 3 * ngc;
        if (ngr > 0) {
            std::vector<int> cs_g(std::max(Ns, 1), -1), rg(std::max(Nr, 1), -1);
            for (int q = 0; q < Ns; q++) {
                const int ro = M->seg_iRiv[c->sperm[cell_seg_idx[q]]] - 1;
                if (ro >= Nr - ngr) cs_g[q] = ro - (Nr - ngr);
            }
            for (int o = Nr - ngr; o < Nr; o++) rg[c->rinv[o]] = o - (Nr - ngr);
            m.cs_g = dev_upload(c, cs_g); m.r_gslot = dev_upload(c, rg);
        }
        if (ngc > 0) {
            std::vector<int> gs(LDh, -1);
            for (int o = Ne - ngc; o < Ne; o++) gs[c->cinv[o]] = o - (Ne - ngc);
            m.g_cslot = dev_upload(c, gs);
        }
        c->n_ghost_cells = ngc; c->n_ghost_reaches = ngr;
    }
    m.bank_cell = dev_upload(c, bank_cell); m.bank_j = dev_upload(c, bank_j); m.bank_lake = dev_upload(c, bank_lake);
    m.bank_kh = dev_upload(c, bank_kh);

    // ---- reaches ----
    m.r_len = up_riv(c, M->riv_Length); m.r_slope = up_riv(c, M->riv_BedSlope); m.r_depth = up_riv(c, M->riv_depth);
    m.r_w0 = up_riv(c, M->riv_BottomWidth); m.r_bank = up_riv(c, M->riv_bankslope);
    m.r_rough = up_riv(c, M->riv_avgRough); m.r_dist = up_riv(c, M->riv_Dist2DownStream);
    m.r_ksatH = up_riv(c, M->riv_KsatH); m.r_bed = up_riv(c, M->riv_BedThick); m.r_zbank = up_riv(c, M->riv_zbank);
    {
        std::vector<int> down(Nr), bc(Nr), toLake(Nr), up_ptr(Nr + 1, 0), up_idx, seg_ptr(Nr + 1, 0);
        std::vector<std::vector<int>> ups(Nr);
        for (int o = 0; o < Nr; o++) {  // ascending reference id => upstream lists come out ascending
            const int r = c->rinv[o];
            const int dn = M->riv_down[o];
            // a partition may hold reaches that flow into a lake held elsewhere (index >= Nl): they keep the to-lake
            // routing formula but feed no local lake
            const int tl = (M->lakeon != 0 && M->riv_toLake) ? M->riv_toLake[o] : -9999;
            bc[r] = M->riv_BC[o];
            toLake[r] = tl >= 0 ? std::min(tl, Nl) : -1;
            if (dn > 0 && dn <= Nr) {
                down[r] = c->rinv[dn - 1];
                // PassValue: iDownStrm >= 0 && toLake <= 0 (MD_f.cpp:237)
                if (tl <= 0) ups[c->rinv[dn - 1]].push_back(r);
            } else {
                down[r] = (dn > 0) ? -99 : (dn == 0 ? -99 : dn);  // outlet codes stay negative; 0 / bad -> error code
                if (dn == 0) down[r] = -99;
            }
        }
        for (int r = 0; r < Nr; r++) {
            up_ptr[r + 1] = up_ptr[r] + (int)ups[r].size();
            for (int u : ups[r]) up_idx.push_back(u);
        }
        for (int s = 0; s < Ns; s++) seg_ptr[c->rinv[M->seg_iRiv[c->sperm[s]] - 1] + 1]++;
        for (int r = 0; r < Nr; r++) seg_ptr[r + 1] += seg_ptr[r];
        m.r_down = dev_upload(c, down); m.r_bc = dev_upload(c, bc); m.r_toLake = dev_upload(c, toLake);
        m.r_up_ptr = dev_upload(c, up_ptr); m.r_up_idx = dev_upload(c, up_idx); m.r_seg_ptr = dev_upload(c, seg_ptr);
        // lakes: inflowing reaches, ascending reference id
        std::vector<int> rin_ptr(Nl + 1, 0), rin_idx;
        for (int l = 0; l < Nl; l++) {
            for (int o = 0; o < Nr; o++)
                if (toLake[c->rinv[o]] == l) rin_idx.push_back(c->rinv[o]);
            rin_ptr[l + 1] = (int)rin_idx.size();
        }
        m.l_rin_ptr = dev_upload(c, rin_ptr); m.l_rin_idx = dev_upload(c, rin_idx);
    }
    // ---- segments ----
    {
        std::vector<int> s_riv(Ns);
        std::vector<double> s_len(Ns), s_cwr(Ns);
        for (int s = 0; s < Ns; s++) {
            const int o = c->sperm[s];
            s_riv[s] = c->rinv[M->seg_iRiv[o] - 1]; s_len[s] = M->seg_length[o]; s_cwr[s] = M->seg_Cwr[o];
        }
        m.s_riv = dev_upload(c, s_riv); m.s_len = dev_upload(c, s_len); m.s_cwr = dev_upload(c, s_cwr);
    }
    // ---- lakes ----
    {
        std::vector<double> zmin(Nl), yi0(Nl), by, ba;
        std::vector<int> bptr(Nl + 1, 0);
        c->lake_nele.assign(Nl, 1);
        for (int l = 0; l < Nl; l++) {
            zmin[l] = M->lake_zmin[l];
            yi0[l] = M->lake_bathy_yi[M->lake_bathy_ptr[l]];
            c->lake_nele[l] = M->lake_NumEleLake[l];
        }
        if (Nl > 0) {
            const int nb = M->lake_bathy_ptr[Nl];
            by.assign(M->lake_bathy_yi, M->lake_bathy_yi + nb);
            ba.assign(M->lake_bathy_ai, M->lake_bathy_ai + nb);
            bptr.assign(M->lake_bathy_ptr, M->lake_bathy_ptr + Nl + 1);
        }
        m.l_zmin = dev_upload(c, zmin); m.l_yi0 = dev_upload(c, yi0); m.l_by = dev_upload(c, by);
        m.l_ba = dev_upload(c, ba); m.l_bptr = dev_upload(c, bptr); m.l_bank_ptr = dev_upload(c, l_bank_ptr);
        m.l_evap_raw = dev_alloc<double>(c, Nl); m.l_prcp = dev_alloc<double>(c, Nl);
        CK(cudaMemset(m.l_evap_raw, 0, sizeof(double) * std::max(Nl, 1)));
        CK(cudaMemset(m.l_prcp, 0, sizeof(double) * std::max(Nl, 1)));
    }
    // ---- halo cells of a partition ----
    m.Nhalo = Nhalo;
    c->Nhalo = Nhalo;
    if (Nhalo) {
        auto uph = [&](const double *src) { return dev_upload(c, std::vector<double>(src, src + Nhalo)); };
        m.h_zs = uph(H->z_surf); m.h_zb = uph(H->z_bottom); m.h_aqd = uph(H->AquiferDepth); m.h_macD = uph(H->macD);
        m.h_macKsatH = uph(H->macKsatH); m.h_vAreaF = uph(H->geo_vAreaF); m.h_ksatH = uph(H->KsatH);
    }
    // ---- dynamic arrays ----
    m.netPrep = dev_alloc<double>(c, LDh); m.potEvap = dev_alloc<double>(c, LDh); m.potTran = dev_alloc<double>(c, LDh);
    m.lai = dev_alloc<double>(c, LDh); m.fuSurf = dev_alloc<double>(c, LDh); m.fuSub = dev_alloc<double>(c, LDh);
    m.eic = dev_alloc<double>(c, LDh); m.satn = dev_alloc<double>(c, LDh);
    m.ele_yBC = dev_alloc<double>(c, LDh); m.ele_QBC = dev_alloc<double>(c, LDh);
    m.r_yBC = dev_alloc<double>(c, Nr); m.r_qBC = dev_alloc<double>(c, Nr);
    m.effKH = dev_alloc<double>(c, LDh); m.QsegSurf = dev_alloc<double>(c, Ns); m.QsegSub = dev_alloc<double>(c, Ns);
    // the error word lives in mapped pinned host memory: kernels raise it with a (rare) system-scope atomic, the host
    // reads it after a stream synchronisation without a copy
    CK(cudaHostAlloc((void **)&c->h_err, sizeof(int) * 2, cudaHostAllocMapped));
    c->h_err[0] = c->h_err[1] = 0;
    CK(cudaHostGetDevicePointer((void **)&m.err, c->h_err, 0));
    for (double *p : {m.netPrep, m.potEvap, m.potTran, m.lai, m.fuSurf, m.fuSub, m.eic, m.satn, m.ele_yBC, m.ele_QBC,
                      m.effKH})
        CK(cudaMemset(p, 0, sizeof(double) * LDh));
    CK(cudaMemset(m.r_yBC, 0, sizeof(double) * std::max(Nr, 1)));
    CK(cudaMemset(m.r_qBC, 0, sizeof(double) * std::max(Nr, 1)));
    c->d_cperm = dev_upload(c, c->cperm); c->d_rperm = dev_upload(c, c->rperm);
    c->y_stage = dev_alloc<double>(c, c->NY); c->y_dev = dev_alloc<double>(c, c->NY);
    c->ydot_dev = dev_alloc<double>(c, c->NY);
    c->h_pinned_n = (size_t)std::max(Ne, Nr);
    CK(cudaMallocHost(&c->h_pinned, sizeof(double) * c->h_pinned_n));
    CK(cudaDeviceSynchronize());
    *out = c;
    return SHUD_OK;
}

void shud_b200_destroy(shud_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->land_stage) cudaFreeHost(c->land_stage);
    for (int k = 0; k < 2; k++)
        if (c->land_ev[k]) cudaEventDestroy(c->land_ev[k]);
    if (c->xstream) cudaStreamSynchronize(c->xstream);
    drop_graphs(c);  // the captured exchange graphs hold NCCL nodes: gone before the communicator
    if (c->nccl_comm && c->nccl_comm_destroy) c->nccl_comm_destroy(c->nccl_comm);
    if (c->nccl_dl) dlclose(c->nccl_dl);
    for (void *q : c->p2p_opened) cudaIpcCloseMemHandle(q);
    if (c->p2p_block) cudaFree(c->p2p_block);
    if (c->xstream) cudaStreamDestroy(c->xstream);
    if (c->ev_pack) cudaEventDestroy(c->ev_pack);
    if (c->ev_kh) cudaEventDestroy(c->ev_kh);
    if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
    for (void *p : c->allocs) cudaFree(p);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_err) cudaFreeHost(c->h_err);
    cudaStreamDestroy(c->stream);
    delete c;
}

int64_t shud_b200_ny(const shud_ctx *c) { return c ? c->NY : 0; }
void *shud_b200_stream(shud_ctx *c) { return c ? (void *)c->stream : nullptr; }
int shud_b200_launches_per_rhs(const shud_ctx *c) { return (c && (c->Nr > 0 || c->Nl > 0)) ? 3 : 2; }

static int upload_perm(shud_ctx *c, double *dst, const double *src, const std::vector<int> &perm) {
    // gather on the host into pinned memory, then one async copy (stream-ordered)
    CK(cudaStreamSynchronize(c->stream));  // h_pinned is reused
    const size_t n = perm.size();
    for (size_t i = 0; i < n; i++) c->h_pinned[i] = src[perm[i]];
    CK(cudaMemcpyAsync(dst, c->h_pinned, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return SHUD_OK;
}

// columns of a forcing upload: reference order in `src` (column k at src + k * ld), device order out
struct PermCols { double *dst[9]; int ncol; };
__global__ void k_perm_cols(const double *__restrict__ src, size_t ld, const int *__restrict__ perm, int n, PermCols p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = perm[i];
    for (int k = 0; k < p.ncol; k++) p.dst[k][i] = src[(size_t)k * ld + o];
}

int shud_b200_set_forcing(shud_ctx *c, const shud_forcing *f) {
    if (!c || !f) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    DevMesh &m = c->m;
    const size_t Ne = (size_t)c->Ne, Nr = (size_t)c->Nr;
    // the arrays go up as they are (reference order) into a device staging area and are permuted there by one kernel
    // per index space: no host gather, no synchronisation per array
    if (!c->f_stage) c->f_stage = dev_alloc<double>(c, 9 * Ne + 2 * std::max<size_t>(Nr, 1));
    const double *src[9] = {f->qEleNetPrep, f->qPotEvap, f->qPotTran, f->t_lai, f->fu_Surf, f->fu_Sub, f->qEleE_IC,
                            f->ele_yBC, f->ele_QBC};
    double *dst[9] = {m.netPrep, m.potEvap, m.potTran, m.lai, m.fuSurf, m.fuSub, m.eic, m.ele_yBC, m.ele_QBC};
    PermCols pc{};
    for (int k = 0; k < 9; k++) {
        if (!src[k]) {
            if (k < 7) return SHUD_ERR_ARG;
            continue;
        }
        CK(cudaMemcpyAsync(c->f_stage + (size_t)pc.ncol * Ne, src[k], sizeof(double) * Ne, cudaMemcpyHostToDevice, c->stream));
        pc.dst[pc.ncol++] = dst[k];
    }
    k_perm_cols<<<(unsigned)((Ne + 255) / 256), 256, 0, c->stream>>>(c->f_stage, Ne, c->d_cperm, (int)Ne, pc);
    if (Nr) {
        PermCols pr{};
        double *rbase = c->f_stage + 9 * Ne;
        const double *rsrc[2] = {f->riv_yBC, f->riv_qBC};
        double *rdst[2] = {m.r_yBC, m.r_qBC};
        for (int k = 0; k < 2; k++) {
            if (!rsrc[k]) continue;
            CK(cudaMemcpyAsync(rbase + (size_t)pr.ncol * Nr, rsrc[k], sizeof(double) * Nr, cudaMemcpyHostToDevice, c->stream));
            pr.dst[pr.ncol++] = rdst[k];
        }
        if (pr.ncol) k_perm_cols<<<(unsigned)((Nr + 255) / 256), 256, 0, c->stream>>>(rbase, Nr, c->d_rperm, (int)Nr, pr);
    }
    CK(cudaGetLastError());
    // lake-cell evaporation / precipitation means, ascending reference cell order (MD_f.cpp:16-17):
    // they depend on the forcing step only (qEleEvapo of a lake cell is qPotEvap, MD_ElementFlux.cpp:15)
    if (c->Nl > 0) {
        if (!f->qElePrep) return SHUD_ERR_ARG;
        std::vector<double> ev(c->Nl, 0.), pr(c->Nl, 0.);
        for (int o : c->lake_cells) {
            const int l = c->lake_of_cell[o];
            ev[l] += f->qPotEvap[o] / c->lake_nele[l];
            pr[l] += f->qElePrep[o] / c->lake_nele[l];
        }
        CK(cudaMemcpyAsync(m.l_evap_raw, ev.data(), sizeof(double) * c->Nl, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(m.l_prcp, pr.data(), sizeof(double) * c->Nl, cudaMemcpyHostToDevice, c->stream));
    }
    // the caller's arrays (pageable host memory, staged by the runtime) may be reused on return
    CK(cudaStreamSynchronize(c->stream));
    return SHUD_OK;
}

int shud_b200_set_carried(shud_ctx *c, const double *satn) {
    if (!c || !satn) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    return upload_perm(c, c->m.satn, satn, c->cperm);
}

int shud_b200_get_carried(shud_ctx *c, double *satn, double *eic) {
    if (!c) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    std::vector<double> h(c->Ne);
    CK(cudaStreamSynchronize(c->stream));
    if (satn) {
        CK(cudaMemcpy(h.data(), c->m.satn, sizeof(double) * c->Ne, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c->Ne; i++) satn[c->cperm[i]] = h[i];
    }
    if (eic) {
        CK(cudaMemcpy(h.data(), c->m.eic, sizeof(double) * c->Ne, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c->Ne; i++) eic[c->cperm[i]] = h[i];
    }
    return SHUD_OK;
}

int shud_b200_to_device_order(shud_ctx *c, const double *ref_dev, double *dev_dev) {
    if (!c) return SHUD_ERR_ARG;
    k_to_dev<<<296, 256, 0, c->stream>>>(ref_dev, dev_dev, c->d_cperm, c->d_rperm, c->Ne, c->Nr, c->Nl);
    CK(cudaGetLastError());
    return SHUD_OK;
}
int shud_b200_from_device_order(shud_ctx *c, const double *dev_dev, double *ref_dev) {
    if (!c) return SHUD_ERR_ARG;
    k_from_dev<<<296, 256, 0, c->stream>>>(dev_dev, ref_dev, c->d_cperm, c->d_rperm, c->Ne, c->Nr, c->Nl);
    CK(cudaGetLastError());
    return SHUD_OK;
}

// host vector in the reference's blocked order <-> device vector in device order (the host mirror of the N_Vector)
int shud_b200_upload_ref(shud_ctx *c, const double *y_host_ref, double *y_dev) {
    if (!c || !y_host_ref || !y_dev) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->y_stage, y_host_ref, sizeof(double) * c->NY, cudaMemcpyHostToDevice, c->stream));
    int rc = shud_b200_to_device_order(c, c->y_stage, y_dev);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));  // y_host_ref may be pageable / reused by the caller
    return SHUD_OK;
}
int shud_b200_download_ref(shud_ctx *c, const double *y_dev, double *y_host_ref) {
    if (!c || !y_host_ref || !y_dev) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    int rc = shud_b200_from_device_order(c, y_dev, c->y_stage);
    if (rc) return rc;
    CK(cudaMemcpyAsync(y_host_ref, c->y_stage, sizeof(double) * c->NY, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SHUD_OK;
}

int shud_b200_set_halo_state_dev(shud_ctx *c, const double *state) {
    if (!c || (c->Nhalo > 0 && !state)) return SHUD_ERR_ARG;
    c->m.h_state = state;
    c->m.h_state_alt = nullptr; c->m.h_epoch = nullptr; c->m.h_flags = nullptr; c->m.h_done = nullptr; c->m.h_nflags = 0;
    c->use_p2p = 0;  // a registered buffer replaces the p2p buffers
    drop_graphs(c);  // kernel parameters changed
    return SHUD_OK;
}

int shud_b200_pack_halo_dev(shud_ctx *c, const double *y, const int32_t *idx, int32_t n, double *out) {
    if (!c || !y || (n > 0 && (!idx || !out))) return SHUD_ERR_ARG;
    if (n <= 0) return SHUD_OK;
    k_pack_halo<<<(n + 255) / 256, 256, 0, c->stream>>>(y, idx, n, c->Ne, out);
    CK(cudaGetLastError());
    return SHUD_OK;
}

// ---- halo exchange over NCCL, driven by the library (one call per f(), no host framework in the step) ----
namespace {
struct nccl_uid { char internal[128]; };  // ncclUniqueId
void *nccl_open(const char *path) {
    void *h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    return h;
}
}  // namespace

int shud_b200_comm_unique_id(const char *nccl_lib, void *id128) {
    if (!id128) return SHUD_ERR_ARG;
    void *h = nccl_open(nccl_lib);
    if (!h) return SHUD_ERR_CUDA;
    auto get = (int (*)(nccl_uid *))dlsym(h, "ncclGetUniqueId");
    if (!get || get((nccl_uid *)id128) != 0) return SHUD_ERR_CUDA;
    return SHUD_OK;
}

int shud_b200_comm_init(shud_ctx *c, const char *nccl_lib, const void *id128, int rank, int world) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->nccl_dl = nccl_open(nccl_lib);
    if (!c->nccl_dl) return SHUD_ERR_CUDA;
    auto init = (int (*)(void **, int, nccl_uid, int))dlsym(c->nccl_dl, "ncclCommInitRank");
    c->nccl_send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))dlsym(c->nccl_dl, "ncclSend");
    c->nccl_recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))dlsym(c->nccl_dl, "ncclRecv");
    c->nccl_group_start = (int (*)())dlsym(c->nccl_dl, "ncclGroupStart");
    c->nccl_group_end = (int (*)())dlsym(c->nccl_dl, "ncclGroupEnd");
    c->nccl_comm_destroy = (int (*)(void *))dlsym(c->nccl_dl, "ncclCommDestroy");
    c->nccl_allreduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(c->nccl_dl, "ncclAllReduce");
    if (!init || !c->nccl_send || !c->nccl_recv || !c->nccl_group_start || !c->nccl_group_end) return SHUD_ERR_CUDA;
    nccl_uid id;
    memcpy(&id, id128, sizeof(id));
    if (init(&c->nccl_comm, world, id, rank) != 0) return SHUD_ERR_CUDA;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
    return SHUD_OK;
}

// Scalar allreduce of the distributed N_Vector reductions ("local partial + ncclAllReduce", batched:
// all the dot products of a Gram-Schmidt sweep travel in one call).  vals: host doubles, reduced in place over the
// ranks of the communicator; op 0 sum, 1 max, 2 min.  Runs on the context stream; synchronises.
constexpr int SHUD_AR_MAX = 64;
int shud_b200_allreduce(shud_ctx *c, double *vals, int n, int op) {
    if (!c || !vals || n < 0 || op < 0 || op > 2) return SHUD_ERR_ARG;
    if (n == 0) return SHUD_OK;
    if (!c->nccl_comm || !c->nccl_allreduce) return SHUD_ERR_ARG;  // shud_b200_comm_init first
    CK(cudaSetDevice(c->device));
    if (!c->ar_dev) c->ar_dev = dev_alloc<double>(c, SHUD_AR_MAX);
    static const int nccl_op[3] = {0 /*ncclSum*/, 2 /*ncclMax*/, 3 /*ncclMin*/};
    for (int k0 = 0; k0 < n; k0 += SHUD_AR_MAX) {
        const int m = std::min(SHUD_AR_MAX, n - k0);
        CK(cudaMemcpyAsync(c->ar_dev, vals + k0, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        if (c->nccl_allreduce(c->ar_dev, c->ar_dev, (size_t)m, 8 /*ncclFloat64*/, nccl_op[op], c->nccl_comm, c->stream) != 0)
            return SHUD_ERR_CUDA;
        CK(cudaMemcpyAsync(vals + k0, c->ar_dev, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return SHUD_OK;
}

// the same on values that are already in device memory, enqueued on `stream` (no copy, no synchronisation): the hook of
// shud_nv_ws_set_allreduce.  `ctx` is the shud_ctx; `stream` must be ordered like the context stream.
int shud_b200_allreduce_dev(void *ctx, double *dev_vals, int n, int op, void *stream) {
    shud_ctx *c = (shud_ctx *)ctx;
    if (!c || !dev_vals || n < 0 || op < 0 || op > 2) return SHUD_ERR_ARG;
    if (n == 0) return SHUD_OK;
    if (!c->nccl_comm || !c->nccl_allreduce) return SHUD_ERR_ARG;
    static const int nccl_op[3] = {0 /*ncclSum*/, 2 /*ncclMax*/, 3 /*ncclMin*/};
    return c->nccl_allreduce(dev_vals, dev_vals, (size_t)n, 8 /*ncclFloat64*/, nccl_op[op], c->nccl_comm, (cudaStream_t)stream) == 0
               ? SHUD_OK : SHUD_ERR_CUDA;
}

int shud_b200_exchange_plan(shud_ctx *c, int npeers, const int32_t *peer_rank, const int32_t *send_count,
                            const int32_t *recv_count, const int32_t *send_cells) {
    if (!c || npeers < 0 || (npeers > 0 && (!peer_rank || !send_count || !recv_count))) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->x_peer.assign(peer_rank, peer_rank + npeers);
    c->x_scount.assign(send_count, send_count + npeers);
    c->x_rcount.assign(recv_count, recv_count + npeers);
    int ns = 0, nr = 0;
    for (int p = 0; p < npeers; p++) { ns += send_count[p]; nr += recv_count[p]; }
    if (nr != c->Nhalo || (ns > 0 && !send_cells)) return SHUD_ERR_ARG;
    std::vector<int> idx(ns);
    for (int k = 0; k < ns; k++) {  // reference-local ids (0-based) -> device order
        if (send_cells[k] < 0 || send_cells[k] >= c->Ne) return SHUD_ERR_ARG;
        idx[k] = c->cinv[send_cells[k]];
    }
    c->x_nsend = ns;
    c->x_scount3.clear(); c->x_rcount3.clear(); c->x_nitems = 0;  // rebuilt from the pair form by shud_b200_p2p_export
    drop_graphs(c);  // captured exchanges point at the previous plan
    if (ns > c->x_cap_send || !c->x_sidx) {  // a repeated plan reuses the buffers of the previous one
        c->x_cap_send = std::max(ns, 1);
        c->x_sidx = dev_alloc<int>(c, (size_t)c->x_cap_send);
        c->x_sbuf = dev_alloc<double>(c, 2 * (size_t)c->x_cap_send);
    }
    if (ns > 0) CK(cudaMemcpy(c->x_sidx, idx.data(), sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice));
    if (!c->x_hstate) c->x_hstate = dev_alloc<double>(c, 2 * (size_t)std::max(nr, 1));  // nr == Nhalo: fixed per context
    CK(cudaMemset(c->x_hstate, 0, sizeof(double) * 2 * (size_t)std::max(nr, 1)));
    return shud_b200_set_halo_state_dev(c, c->x_hstate);
}

// ---- peer-to-peer exchange: set-up ----
struct P2PBlob {
    cudaIpcMemHandle_t handle;   // 64 bytes
    long long pid;               // contexts of one process are connected by pointer
    void *base;
    int rank, npeers;
    unsigned long long stride;
    int peer[P2P_MAXPEER], recv_off[P2P_MAXPEER][3];  // where neighbour p's doubles of each kind land in my halo buffer
};
static_assert(sizeof(P2PBlob) <= SHUD_P2P_BLOB_BYTES, "blob too large");
constexpr size_t P2P_HDR = 256;  // flags [P2P_MAXPEER] | epoch | count

// device-order flat index of entry `k` (reference-local blocked order) of this partition's state vector
static int flat_to_device(const shud_ctx *c, int k) {
    const int Ne = c->Ne, Nr = c->Nr;
    if (k < 0 || k >= c->NY) return -1;
    if (k < 3 * Ne) return (k / Ne) * Ne + c->cinv[k % Ne];
    if (k < 3 * Ne + Nr) return 3 * Ne + c->rinv[k - 3 * Ne];
    return k;
}

int shud_b200_exchange_plan_items(shud_ctx *c, int npeers, const int32_t *peer_rank, const int32_t *send_count,
                                  const int32_t *recv_count, const int32_t *send_items) {
    if (!c || npeers < 0 || npeers > P2P_MAXPEER || (npeers > 0 && (!peer_rank || !send_count || !recv_count))) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    int ns = 0, nr[3] = {0, 0, 0};
    for (int p = 0; p < npeers; p++)
        for (int k = 0; k < 3; k++) {
            if (send_count[3 * p + k] < 0 || recv_count[3 * p + k] < 0) return SHUD_ERR_ARG;
            ns += send_count[3 * p + k]; nr[k] += recv_count[3 * p + k];
        }
    if (nr[0] != 2 * c->Nhalo || nr[1] != 3 * c->n_ghost_cells || nr[2] != c->n_ghost_reaches || (ns > 0 && !send_items))
        return SHUD_ERR_ARG;
    std::vector<int> idx(std::max(ns, 1), 0);
    for (int k = 0; k < ns; k++) {
        idx[k] = flat_to_device(c, send_items[k]);
        if (idx[k] < 0) return SHUD_ERR_ARG;
    }
    drop_graphs(c);
    c->x_peer.assign(peer_rank, peer_rank + npeers);
    c->x_scount3.assign(send_count, send_count + 3 * npeers);
    c->x_rcount3.assign(recv_count, recv_count + 3 * npeers);
    c->x_nitems = ns;
    c->x_items = dev_upload(c, idx);
    // the pair form of the plan (NCCL path, shud_b200_exchange_plan) only exists without ghosts
    c->x_scount.assign(npeers, 0); c->x_rcount.assign(npeers, 0);
    for (int p = 0; p < npeers; p++) { c->x_scount[p] = send_count[3 * p] / 2; c->x_rcount[p] = recv_count[3 * p] / 2; }
    c->use_p2p = 0;
    return SHUD_OK;
}

int shud_b200_p2p_export(shud_ctx *c, int rank, void *blob) {
    if (!c || !blob || (int)c->x_peer.size() > P2P_MAXPEER) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const int npeers = (int)c->x_peer.size();
    if (c->x_rcount3.empty() && npeers > 0) {
        // plan given in the pair form (shud_b200_exchange_plan): halo pairs only
        if (c->n_ghost_cells || c->n_ghost_reaches) return SHUD_ERR_ARG;
        std::vector<int> items;
        c->x_scount3.assign(3 * npeers, 0); c->x_rcount3.assign(3 * npeers, 0);
        std::vector<int> sidx(std::max(c->x_nsend, 1));
        if (c->x_nsend) CK(cudaMemcpy(sidx.data(), c->x_sidx, sizeof(int) * c->x_nsend, cudaMemcpyDeviceToHost));
        for (int k = 0; k < c->x_nsend; k++) { items.push_back(sidx[k]); items.push_back(2 * c->Ne + sidx[k]); }
        for (int p = 0; p < npeers; p++) { c->x_scount3[3 * p] = 2 * c->x_scount[p]; c->x_rcount3[3 * p] = 2 * c->x_rcount[p]; }
        c->x_nitems = (int)items.size();
        if (items.empty()) items.push_back(0);
        c->x_items = dev_upload(c, items);
    }
    const size_t ndbl = (size_t)2 * c->Nhalo + 3 * (size_t)c->n_ghost_cells + (size_t)c->n_ghost_reaches;
    if (!c->p2p_block) {
        c->p2p_stride = (std::max<size_t>(ndbl, 1) * sizeof(double) + 255) / 256 * 256;
        // its own allocation: one IPC handle, nothing else exposed; behind the two halo buffers the mailbox of the
        // in-kernel allreduce of the distributed vector's reductions (shud_nv_ws_set_peer_allreduce)
        CK(cudaMalloc(&c->p2p_block, P2P_HDR + 2 * c->p2p_stride + SHUD_NV_ARBOX_BYTES));
        CK(cudaMemset(c->p2p_block, 0, P2P_HDR + 2 * c->p2p_stride + SHUD_NV_ARBOX_BYTES));
    }
    P2PBlob b;
    memset(&b, 0, sizeof(b));
    CK(cudaIpcGetMemHandle(&b.handle, c->p2p_block));
    b.pid = (long long)getpid(); b.base = c->p2p_block;
    b.rank = rank; b.npeers = npeers; b.stride = c->p2p_stride;
    // a kind's region is filled neighbour by neighbour, in the order the halo cells / ghosts are numbered
    int ro[3] = {0, 2 * c->Nhalo, 2 * c->Nhalo + 3 * c->n_ghost_cells};
    for (int p = 0; p < npeers; p++) {
        b.peer[p] = c->x_peer[p];
        for (int k = 0; k < 3; k++) { b.recv_off[p][k] = ro[k]; ro[k] += c->x_rcount3[3 * p + k]; }
    }
    memset(blob, 0, SHUD_P2P_BLOB_BYTES);
    memcpy(blob, &b, sizeof(b));
    return SHUD_OK;
}

int shud_b200_p2p_mailboxes(shud_ctx *c, int *nranks, int *rank, void **boxes) {
    if (!c || !nranks || !rank || !boxes) return SHUD_ERR_ARG;
    *nranks = c->ar_nranks; *rank = c->ar_rank;
    for (int r = 0; r < c->ar_nranks; r++) boxes[r] = c->ar_box[r];
    return SHUD_OK;
}

int shud_b200_p2p_connect(shud_ctx *c, int rank, int world, const void *blobs) {
    if (!c || !blobs || !c->p2p_block || rank < 0 || rank >= world) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const char *env = getenv("SHUD_P2P");
    if (env && atoi(env) == 0) return SHUD_OK;  // keep the NCCL path (A/B)
    for (void *q : c->p2p_opened) cudaIpcCloseMemHandle(q);  // a repeated connect maps afresh
    c->p2p_opened.clear();
    c->use_p2p = 0;
    P2PTable T{};
    T.npeers = (int)c->x_peer.size();
    int so = 0;
    c->ar_nranks = 0;
    // my mailbox starts from zero tags with every (re)connect: a workspace that installs these mailboxes counts from 1
    // (no peer stores into it before its own connect and the barrier the callers hold after this call)
    CK(cudaMemset((char *)c->p2p_block + P2P_HDR + 2 * c->p2p_stride, 0, SHUD_NV_ARBOX_BYTES));
    std::vector<char *> mapped(world, nullptr);  // block of rank r as this process sees it
    mapped[rank] = (char *)c->p2p_block;
    for (int p = 0; p < T.npeers; p++) {
        const int r = c->x_peer[p];
        if (r < 0 || r >= world) return SHUD_ERR_ARG;
        P2PBlob b;
        memcpy(&b, (const char *)blobs + (size_t)r * SHUD_P2P_BLOB_BYTES, sizeof(b));
        int slot = -1;
        for (int j = 0; j < b.npeers; j++) if (b.peer[j] == rank) slot = j;
        if (b.rank != r || slot < 0) return SHUD_ERR_ARG;
        char *base = nullptr;
        if (r == rank) {
            base = (char *)c->p2p_block;
        } else if (b.pid == (long long)getpid()) {
            base = (char *)b.base;  // same process: the pointer itself (IPC handles cannot be opened by their creator)
            int dev_peer = -1;
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, base) == cudaSuccess) dev_peer = pa.device;
            if (dev_peer >= 0 && dev_peer != c->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(dev_peer, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SHUD_ERR_CUDA; }
                cudaGetLastError();
            }
        } else {
            if (cudaIpcOpenMemHandle((void **)&base, b.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                for (void *q : c->p2p_opened) cudaIpcCloseMemHandle(q);
                c->p2p_opened.clear();
                return SHUD_ERR_CUDA;  // caller falls back to the NCCL path
            }
            c->p2p_opened.push_back(base);
        }
        mapped[r] = base;
        T.buf[0][p] = (double *)(base + P2P_HDR);
        T.buf[1][p] = (double *)(base + P2P_HDR + b.stride);
        T.flag[p] = (unsigned long long *)base + slot;
        for (int k = 0; k < 3; k++) {
            const int n = c->x_scount3[3 * p + k];
            if (n == 0) continue;
            T.seg_start[T.nseg] = so; T.seg_peer[T.nseg] = p; T.seg_dst[T.nseg] = b.recv_off[slot][k];
            T.nseg++; so += n;
        }
    }
    T.seg_start[T.nseg] = so;
    if (so != c->x_nitems) return SHUD_ERR_ARG;
    c->p2p = T;
    // my own side: the two halo buffers, the flags, the epoch word
    char *mine = (char *)c->p2p_block;
    c->m.h_state = (const double *)(mine + P2P_HDR);
    c->m.h_state_alt = (const double *)(mine + P2P_HDR + c->p2p_stride);
    c->m.h_epoch = (unsigned long long *)mine + P2P_MAXPEER;
    c->m.h_flags = (const unsigned long long *)mine;
    c->m.h_done = (unsigned int *)((unsigned long long *)mine + P2P_MAXPEER + 1) + 1;  // beside the pack counter
    c->m.h_nflags = T.npeers;
    c->m.n_int_tiles = c->n_int_tiles;
    // mailboxes of the in-kernel allreduce: every rank's block, not only the neighbours' (other processes only; a rank
    // that cannot be mapped leaves the allreduce to NCCL)
    if (world >= 2 && world <= SHUD_NV_MAXRANKS) {
        bool ok = true;
        for (int r = 0; r < world && ok; r++) {
            P2PBlob b;
            memcpy(&b, (const char *)blobs + (size_t)r * SHUD_P2P_BLOB_BYTES, sizeof(b));
            if (b.rank != r) { ok = false; break; }
            if (!mapped[r]) {
                if (b.pid == (long long)getpid()) { ok = false; break; }  // several contexts of one process: no mailbox allreduce
                char *base = nullptr;
                if (cudaIpcOpenMemHandle((void **)&base, b.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                c->p2p_opened.push_back(base);
                mapped[r] = base;
            }
            c->ar_box[r] = mapped[r] + P2P_HDR + 2 * b.stride;
        }
        if (ok) { c->ar_nranks = world; c->ar_rank = rank; }
    }
    drop_graphs(c);
    c->use_p2p = 1;
    return SHUD_OK;
}

static int exchange_launch(shud_ctx *c, double t, const double *y, double *ydot) {
    // pack my boundary cells -> post the sends / receives on the exchange stream -> interior part of f() beside
    // them on the context stream -> boundary part (its tiles on the exchange stream, behind the receives)
    if (c->x_nsend > 0)
        k_pack_halo<<<(c->x_nsend + 255) / 256, 256, 0, c->stream>>>(y, c->x_sidx, c->x_nsend, c->Ne, c->x_sbuf);
    CK(cudaEventRecord(c->ev_pack, c->stream));
    CK(cudaStreamWaitEvent(c->xstream, c->ev_pack, 0));
    if (!c->x_peer.empty()) {
        const int f64 = 8;  // ncclFloat64
        if (c->nccl_group_start() != 0) return SHUD_ERR_CUDA;
        size_t so = 0, ro = 0;
        for (size_t p = 0; p < c->x_peer.size(); p++) {
            const size_t ns = 2 * (size_t)c->x_scount[p], nr = 2 * (size_t)c->x_rcount[p];
            if (ns && c->nccl_send(c->x_sbuf + so, ns, f64, c->x_peer[p], c->nccl_comm, c->xstream) != 0) return SHUD_ERR_CUDA;
            if (nr && c->nccl_recv(c->x_hstate + ro, nr, f64, c->x_peer[p], c->nccl_comm, c->xstream) != 0) return SHUD_ERR_CUDA;
            so += ns; ro += nr;
        }
        if (c->nccl_group_end() != 0) return SHUD_ERR_CUDA;
    }
    int rc = shud_b200_rhs_interior_dev(c, t, y, ydot);
    if (rc) return rc;
    return shud_b200_rhs_boundary_dev(c, t, y, ydot, c->xstream);
}

int shud_b200_rhs_exchange_dev(shud_ctx *c, double t, const double *y, double *ydot) {
    if (!c || !y || !ydot) return SHUD_ERR_ARG;
    if (c->use_p2p) return shud_b200_rhs_dev(c, t, y, ydot);  // peer-to-peer: every f() of this context exchanges
    if (!c->nccl_comm || !c->xstream) return SHUD_ERR_ARG;    // shud_b200_comm_init + shud_b200_exchange_plan first
    if (!c->x_sidx && c->Nhalo > 0) return SHUD_ERR_ARG;      // an item plan (shud_b200_exchange_plan_items) needs the peer-to-peer transport
    if (!c->use_xgraph) return exchange_launch(c, t, y, ydot);
    // as shud_b200_rhs_dev: the sequence (collective included) is fixed, one instantiated graph per pointer pair
    for (auto &g : c->xgraphs)
        if (g.y == y && g.yd == ydot) {
            g.used = ++c->graph_clock;
            CK(cudaGraphLaunch(g.exec, c->stream));
            return SHUD_OK;
        }
    evict_lru(c, c->xgraphs);
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->use_xgraph = 0;
        return exchange_launch(c, t, y, ydot);
    }
    int rc = exchange_launch(c, t, y, ydot);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == SHUD_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != SHUD_OK || e != cudaSuccess || !exec) {
        cudaGetLastError();
        c->use_xgraph = 0;  // capture of the collective unavailable: plain launches
        return exchange_launch(c, t, y, ydot);
    }
    c->xgraphs.push_back({y, ydot, exec, ++c->graph_clock});
    CK(cudaGraphLaunch(exec, c->stream));
    return SHUD_OK;
}

int shud_b200_perm(const shud_ctx *c, int32_t *cp, int32_t *rp) {
    if (!c) return SHUD_ERR_ARG;
    if (cp) std::copy(c->cperm.begin(), c->cperm.end(), cp);
    if (rp) std::copy(c->rperm.begin(), c->rperm.end(), rp);
    return SHUD_OK;
}

int shud_b200_summary_dev(shud_ctx *c, const double *y_dev, double *y_host_ref) {
    if (!c || !y_dev || !y_host_ref) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    k_summary<<<(unsigned)((c->NY + 255) / 256), 256, 0, c->stream>>>(c->m, y_dev, c->ydot_dev, (size_t)c->NY);
    CK(cudaGetLastError());
    int rc = shud_b200_from_device_order(c, c->ydot_dev, c->y_stage);
    if (rc) return rc;
    CK(cudaMemcpyAsync(y_host_ref, c->y_stage, sizeof(double) * c->NY, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SHUD_OK;
}

int shud_b200_prime(shud_ctx *c, const double *y_host) {
    if (!c || !y_host) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->y_stage, y_host, sizeof(double) * c->NY, cudaMemcpyHostToDevice, c->stream));
    int rc = shud_b200_to_device_order(c, c->y_stage, c->y_dev);
    if (rc) return rc;
    k_prime<<<(c->Ne + 255) / 256, 256, 0, c->stream>>>(c->m, c->y_dev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return SHUD_OK;
}

static int ensure_diag(shud_ctx *c) {
    if (c->diag_alloc) return SHUD_OK;
    DevDiag &d = c->diag;
    const int Ne = c->Ne, Nr = c->Nr, Nl = c->Nl;
    double **cell[] = {&d.qEleInfil, &d.qEleExfil, &d.qEleRecharge, &d.qEs, &d.qEu, &d.qEg, &d.qTu, &d.qTg,
                       &d.qEleTrans, &d.qEleEvapo, &d.qEleETA, &d.iBeta, &d.QeleSurfTot, &d.QeleSubTot, &d.Qe2r_Surf,
                       &d.Qe2r_Sub};
    for (double **p : cell) {
        *p = dev_alloc<double>(c, Ne);
        if (!*p) return SHUD_ERR_CUDA;
        cudaMemset(*p, 0, sizeof(double) * Ne);
    }
    d.QeleSurf = dev_alloc<double>(c, 3 * (size_t)Ne); d.QeleSub = dev_alloc<double>(c, 3 * (size_t)Ne);
    double **riv[] = {&d.QrivSurf, &d.QrivSub, &d.QrivUp, &d.QrivDown};
    for (double **p : riv) *p = dev_alloc<double>(c, Nr);
    double **lake[] = {&d.y2LakeArea, &d.QLakeSurf, &d.QLakeSub, &d.QLakeRivIn, &d.QLakeRivOut, &d.qLakeEvap,
                       &d.qLakePrcp};
    for (double **p : lake) *p = dev_alloc<double>(c, Nl);
    c->diag_alloc = true;
    return SHUD_OK;
}

}  // extern "C"
// a context with halo cells or ghosts runs the kernels compiled with the exchange code (HALO = true)
static inline bool is_partition(const shud_ctx *c) {
    return c->Nhalo > 0 || c->n_ghost_cells > 0 || c->n_ghost_reaches > 0 || c->use_p2p || c->force_halo;
}
// the pre-pass; on a partition connected peer-to-peer it carries the send side of the halo exchange, so EVERY f() of such
// a context exchanges (all ranks make the same calls)
static void launch_prepass(shud_ctx *c, const double *y) {
    const int nb = (c->Ne + 255) / 256;
    if (c->use_p2p) {
        PackArgs P;
        P.T = c->p2p; P.idx = c->x_items; P.n = c->x_nitems;
        P.nblk = std::min(nb, std::max(1, (c->x_nitems + 255) / 256));
        P.count = (unsigned int *)((unsigned long long *)c->p2p_block + P2P_MAXPEER + 1);
        k_effkh_pack<<<nb, 256, 0, c->stream>>>(c->m, y, P);
    } else {
        k_effkh<<<nb, 256, 0, c->stream>>>(c->m, y);
    }
}
// which instantiation of the cell / river kernels a context runs: 0 single domain, 1 partition with halo cells only,
// 2 partition with ghost cells / reaches (cut river trees)
static inline int halo_level(const shud_ctx *c) {
    if (c->n_ghost_cells > 0 || c->n_ghost_reaches > 0 || c->force_halo == 2) return 2;
    return is_partition(c) ? 1 : 0;
}
// one launch of the cell kernel over tiles [tile0, tile0 + ntiles) on `stream`, optionally programmatically dependent
template <bool DIAG>
static cudaError_t launch_cells(shud_ctx *c, int ntiles, int tile0, cudaStream_t stream, const double *y, double *ydot, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles); cfg.blockDim = dim3(2 * TILE); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    switch (halo_level(c)) {
        case 0: return cudaLaunchKernelEx(&cfg, k_fused<DIAG, 4, 0>, c->m, c->diag, y, ydot, tile0);
        case 1: return cudaLaunchKernelEx(&cfg, k_fused<DIAG, 4, 1>, c->m, c->diag, y, ydot, tile0);
        default: return cudaLaunchKernelEx(&cfg, k_fused<DIAG, 4, 2>, c->m, c->diag, y, ydot, tile0);
    }
}
template <bool DIAG>
static cudaError_t launch_river(shud_ctx *c, int nblocks, int nb_riv, const double *y, double *ydot, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblocks); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    switch (halo_level(c)) {
        case 0: return cudaLaunchKernelEx(&cfg, k_river_lake<DIAG, 0>, c->m, c->diag, y, ydot, nb_riv);
        case 1: return cudaLaunchKernelEx(&cfg, k_river_lake<DIAG, 1>, c->m, c->diag, y, ydot, nb_riv);
        default: return cudaLaunchKernelEx(&cfg, k_river_lake<DIAG, 2>, c->m, c->diag, y, ydot, nb_riv);
    }
}
template <bool DIAG>
static void launch_fused(shud_ctx *c, const double *y, double *ydot, bool pdl = false) {
    const int nb = (c->Ne + TILE - 1) / TILE;
    if (pdl && c->use_pdl) {
        // launched while k_effkh is still running (programmatic stream serialisation)
        if (launch_cells<DIAG>(c, nb, 0, c->stream, y, ydot, true) == cudaSuccess) return;
        cudaGetLastError();
        c->use_pdl = 0;
    }
    launch_cells<DIAG>(c, nb, 0, c->stream, y, ydot, false);
}
template <bool DIAG>
static int launch_rhs(shud_ctx *c, const double *y, double *ydot) {
    launch_prepass(c, y);
    launch_fused<DIAG>(c, y, ydot, true);
    const int nb_riv = (c->Nr + 127) / 128;
    if (nb_riv + c->Nl == 0 && c->use_p2p) {  // no reach, no lake: a one-block launch that only advances the epoch word
        launch_river<DIAG>(c, 1, 1, y, ydot, false);
    }
    if (nb_riv + c->Nl > 0) {
        bool done = false;
        if (c->use_pdl) {
            done = launch_river<DIAG>(c, nb_riv + c->Nl, nb_riv, y, ydot, true) == cudaSuccess;
            if (!done) { cudaGetLastError(); c->use_pdl = 0; }
        }
        if (!done) launch_river<DIAG>(c, nb_riv + c->Nl, nb_riv, y, ydot, false);
    }
    CK(cudaGetLastError());
    return SHUD_OK;
}
extern "C" {

static void drop_graphs(shud_ctx *c) {
    for (auto &g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    for (auto &g : c->xgraphs) cudaGraphExecDestroy(g.exec);
    c->xgraphs.clear();
    for (auto &g : c->dqgraphs) cudaGraphExecDestroy(g.exec);
    c->dqgraphs.clear();
}

// f(t, y0 + sigma v ./ ewt) with the perturbation formed by the pre-pass (k_effkh_dq); ytemp receives the perturbed
// state.  A single domain only: a partition's pre-pass carries the halo exchange (the caller perturbs with
// shud_nv_dq_perturb and calls shud_b200_rhs_dev / _rhs_exchange_dev there).
static int launch_rhs_dq(shud_ctx *c, const DqArgs &A, double *ydot) {
    k_effkh_dq<<<(c->Ne + 255) / 256, 256, 0, c->stream>>>(c->m, A);
    // no programmatic launch of the cell kernel here: its vertical role reads ytemp from its first instruction on
    launch_fused<false>(c, A.yt, ydot, false);
    const int nb_riv = (c->Nr + 127) / 128;
    if (nb_riv + c->Nl > 0) {
        bool done = false;
        if (c->use_pdl) {
            done = launch_river<false>(c, nb_riv + c->Nl, nb_riv, A.yt, ydot, true) == cudaSuccess;
            if (!done) { cudaGetLastError(); c->use_pdl = 0; }
        }
        if (!done) launch_river<false>(c, nb_riv + c->Nl, nb_riv, A.yt, ydot, false);
    }
    CK(cudaGetLastError());
    return SHUD_OK;
}
int shud_b200_dq_foldable(const shud_ctx *c) { return c && halo_level(c) == 0 && !c->use_p2p ? 1 : 0; }

int shud_b200_rhs_dq_dev(shud_ctx *c, double t, double sigma, const double *v, const double *ewt, const double *y0,
                         double *ytemp, double *ydot, const double *ss, double *v_out) {
    (void)t;
    if (!c || !v || !ewt || !y0 || !ytemp || !ydot || (v_out && v_out == v)) return SHUD_ERR_ARG;
    if (halo_level(c) != 0 || c->use_p2p) return SHUD_ERR_ARG;
    DqArgs A;
    A.sigma = sigma; A.v = v; A.ewt = ewt; A.y0 = y0; A.yt = ytemp; A.ss = ss; A.v_out = v_out;
    if (!c->use_graph) return launch_rhs_dq(c, A, ydot);
    for (auto &g : c->dqgraphs)
        if (g.a.v == v && g.yd == ydot && g.a.ewt == ewt && g.a.y0 == y0 && g.a.yt == ytemp && g.a.sigma == sigma &&
            g.a.ss == ss && g.a.v_out == v_out) {
            g.used = ++c->graph_clock;
            CK(cudaGraphLaunch(g.exec, c->stream));
            return SHUD_OK;
        }
    if (c->dqgraphs.size() >= 32) {
        size_t lru = 0;
        for (size_t k = 1; k < c->dqgraphs.size(); k++)
            if (c->dqgraphs[k].used < c->dqgraphs[lru].used) lru = k;
        cudaGraphExecDestroy(c->dqgraphs[lru].exec);
        c->dqgraphs.erase(c->dqgraphs.begin() + lru);
    }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = launch_rhs_dq(c, A, ydot);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc == SHUD_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != SHUD_OK || e != cudaSuccess || !exec) {
        cudaGetLastError();
        return launch_rhs_dq(c, A, ydot);
    }
    c->dqgraphs.push_back({A, ydot, exec, ++c->graph_clock});
    CK(cudaGraphLaunch(exec, c->stream));
    return SHUD_OK;
}

int shud_b200_rhs_dev(shud_ctx *c, double t, const double *y, double *ydot) {
    (void)t;  // f() depends on t only through values uploaded by shud_b200_set_forcing
    if (!c || !y || !ydot) return SHUD_ERR_ARG;
    // a partition needs somewhere to read its halo / ghost states from: a registered buffer or the peer-to-peer buffers
    if ((c->Nhalo > 0 || c->n_ghost_cells > 0 || c->n_ghost_reaches > 0) && !c->m.h_state) return SHUD_ERR_ARG;
    if (!c->use_graph) return launch_rhs<false>(c, y, ydot);
    // the launch sequence is fixed; only the two vector pointers vary between calls (CVODE alternates
    // between a handful of work vectors) -> one instantiated graph per pointer pair, launched as one unit
    for (auto &g : c->graphs)
        if (g.y == y && g.yd == ydot) {
            g.used = ++c->graph_clock;
            CK(cudaGraphLaunch(g.exec, c->stream));
            return SHUD_OK;
        }
    evict_lru(c, c->graphs);
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = launch_rhs<false>(c, y, ydot);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != SHUD_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        c->use_graph = 0;  // capture unavailable: plain launches
        return launch_rhs<false>(c, y, ydot);
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        c->use_graph = 0;
        return launch_rhs<false>(c, y, ydot);
    }
    c->graphs.push_back({y, ydot, exec, ++c->graph_clock});
    CK(cudaGraphLaunch(exec, c->stream));
    return SHUD_OK;
}

// The RHS of a partition in two parts so that the halo exchange overlaps the bulk of the work:
//   interior: effKH of the owned cells + the cell kernel on every tile that sees no halo cell (needs no exchanged data)
//   boundary: effKH of the halo cells + the cell kernel on the remaining tiles + the river/lake kernel
int shud_b200_rhs_interior_dev(shud_ctx *c, double t, const double *y, double *ydot) {
    (void)t;
    if (!c || !y || !ydot) return SHUD_ERR_ARG;
    k_effkh<<<(c->Ne + 255) / 256, 256, 0, c->stream>>>(c->m, y);
    if (!c->ev_kh) {
        CK(cudaEventCreateWithFlags(&c->ev_kh, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
    }
    CK(cudaEventRecord(c->ev_kh, c->stream));
    if (c->n_int_tiles > 0)
        launch_cells<false>(c, c->n_int_tiles, 0, c->stream, y, ydot, false);
    CK(cudaGetLastError());
    return SHUD_OK;
}
int shud_b200_rhs_boundary_dev(shud_ctx *c, double t, const double *y, double *ydot, void *halo_stream) {
    (void)t;
    if (!c || !y || !ydot) return SHUD_ERR_ARG;
    // the halo-dependent tiles go on the stream the exchange completes on: they run beside the interior tiles of
    // the context stream (a few hundred blocks in the gaps of ~8000) instead of as a short serial pass behind them
    cudaStream_t hs = halo_stream ? (cudaStream_t)halo_stream : c->stream;
    const bool side = hs != c->stream && c->ev_kh;
    if (side) CK(cudaStreamWaitEvent(hs, c->ev_kh, 0));
    if (c->n_bnd_tiles > 0)
        launch_cells<false>(c, c->n_bnd_tiles, c->n_int_tiles, hs, y, ydot, false);
    if (side) {
        CK(cudaEventRecord(c->ev_bnd, hs));
        CK(cudaStreamWaitEvent(c->stream, c->ev_bnd, 0));
    }
    const int nb_riv = (c->Nr + 127) / 128;
    if (nb_riv + c->Nl > 0) launch_river<false>(c, nb_riv + c->Nl, nb_riv, y, ydot, false);
    CK(cudaGetLastError());
    return SHUD_OK;
}

int shud_b200_tile_counts(const shud_ctx *c, int *n_interior, int *n_boundary) {
    if (!c) return SHUD_ERR_ARG;
    if (n_interior) *n_interior = c->n_int_tiles;
    if (n_boundary) *n_boundary = c->n_bnd_tiles;
    return SHUD_OK;
}

// one launch of the sequence on its own (profiling / per-kernel CUDA-event timing in bench.py)
int shud_b200_rhs_stage_dev(shud_ctx *c, int stage, const double *y, double *ydot) {
    if (!c || !y || !ydot) return SHUD_ERR_ARG;
    const int nb_riv = (c->Nr + 127) / 128;
    if (stage == 0) launch_prepass(c, y);
    else if (stage == 1) launch_fused<false>(c, y, ydot);
    else if (stage == 2 && nb_riv + c->Nl > 0) {
        launch_river<false>(c, nb_riv + c->Nl, nb_riv, y, ydot, false);
    } else return SHUD_ERR_ARG;
    CK(cudaGetLastError());
    return SHUD_OK;
}

int shud_b200_rhs_diag_dev(shud_ctx *c, double t, const double *y, double *ydot) {
    (void)t;
    if (!c || !y || !ydot) return SHUD_ERR_ARG;
    if ((c->Nhalo > 0 || c->n_ghost_cells > 0 || c->n_ghost_reaches > 0) && !c->m.h_state) return SHUD_ERR_ARG;
    int rc = ensure_diag(c);
    if (rc) return rc;
    return launch_rhs<true>(c, y, ydot);
}

int shud_b200_check(shud_ctx *c, int32_t *where) {
    if (!c) return SHUD_ERR_ARG;
    CK(cudaStreamSynchronize(c->stream));
    const int h[2] = {((volatile int *)c->h_err)[0], ((volatile int *)c->h_err)[1]};
    if (h[0]) {
        c->h_err[0] = c->h_err[1] = 0;
        // device ids -> reference ids (code 1 is raised by reaches, the others by cells)
        if (where) {
            const int k = h[1] - 1;
            if (h[0] == SHUD_ERRRIVBC) *where = (k >= 0 && k < c->Nr) ? c->rperm[k] + 1 : 0;
            else *where = (k >= 0 && k < c->Ne) ? c->cperm[k] + 1 : 0;
        }
    } else if (where) {
        *where = 0;
    }
    return h[0];
}

int shud_b200_rhs(shud_ctx *c, double t, const double *y_host, double *ydot_host) {
    if (!c || !y_host || !ydot_host) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->y_stage, y_host, sizeof(double) * c->NY, cudaMemcpyHostToDevice, c->stream));
    int rc = shud_b200_to_device_order(c, c->y_stage, c->y_dev);
    if (rc) return rc;
    rc = shud_b200_rhs_dev(c, t, c->y_dev, c->ydot_dev);
    if (rc) return rc;
    rc = shud_b200_from_device_order(c, c->ydot_dev, c->y_stage);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ydot_host, c->y_stage, sizeof(double) * c->NY, cudaMemcpyDeviceToHost, c->stream));
    return shud_b200_check(c, nullptr);
}

static int download_diag(shud_ctx *c, const DevDiag &d, const double *effKH, const double *satn, const double *qsegS,
                         const double *qsegG, const shud_diag *o);

// the (array, accumulator, length) triples of the output accumulation
static void acc_list(shud_ctx *c, std::vector<const double *> &src, std::vector<double *> &dst, std::vector<size_t> &len) {
    const size_t Ne = c->Ne, Nr = c->Nr, Ns = c->Ns, Nl = c->Nl;
    const DevDiag &d = c->diag;
    DevDiag &a = c->acc;
    auto add = [&](const double *s_, double *d_, size_t n_) { src.push_back(s_); dst.push_back(d_); len.push_back(n_); };
    add(d.qEleInfil, a.qEleInfil, Ne); add(d.qEleExfil, a.qEleExfil, Ne); add(d.qEleRecharge, a.qEleRecharge, Ne);
    add(d.qEs, a.qEs, Ne); add(d.qEu, a.qEu, Ne); add(d.qEg, a.qEg, Ne); add(d.qTu, a.qTu, Ne); add(d.qTg, a.qTg, Ne);
    add(d.qEleTrans, a.qEleTrans, Ne); add(d.qEleEvapo, a.qEleEvapo, Ne); add(d.qEleETA, a.qEleETA, Ne);
    add(d.iBeta, a.iBeta, Ne); add(d.QeleSurf, a.QeleSurf, 3 * Ne); add(d.QeleSub, a.QeleSub, 3 * Ne);
    add(d.QeleSurfTot, a.QeleSurfTot, Ne); add(d.QeleSubTot, a.QeleSubTot, Ne); add(d.Qe2r_Surf, a.Qe2r_Surf, Ne);
    add(d.Qe2r_Sub, a.Qe2r_Sub, Ne); add(d.QrivSurf, a.QrivSurf, Nr); add(d.QrivSub, a.QrivSub, Nr);
    add(d.QrivUp, a.QrivUp, Nr); add(d.QrivDown, a.QrivDown, Nr); add(d.y2LakeArea, a.y2LakeArea, Nl);
    add(d.QLakeSurf, a.QLakeSurf, Nl); add(d.QLakeSub, a.QLakeSub, Nl); add(d.QLakeRivIn, a.QLakeRivIn, Nl);
    add(d.QLakeRivOut, a.QLakeRivOut, Nl); add(d.qLakeEvap, a.qLakeEvap, Nl); add(d.qLakePrcp, a.qLakePrcp, Nl);
    add(c->m.effKH, c->acc_effKH, Ne); add(c->m.satn, c->acc_satn, Ne);
    add(c->m.QsegSurf, c->acc_QsegSurf, Ns); add(c->m.QsegSub, c->acc_QsegSub, Ns);
}

int shud_b200_output_accumulate(shud_ctx *c) {
    if (!c || !c->diag_alloc) return SHUD_ERR_ARG;  // needs a shud_b200_rhs_diag_dev before
    CK(cudaSetDevice(c->device));
    if (!c->acc_alloc) {
        const size_t Ne = c->Ne, Nr = c->Nr, Ns = c->Ns, Nl = c->Nl;
        DevDiag &a = c->acc;
        auto z = [&](size_t n_) { double *p_ = dev_alloc<double>(c, n_); if (p_) cudaMemset(p_, 0, sizeof(double) * std::max<size_t>(n_, 1)); return p_; };
        a.qEleInfil = z(Ne); a.qEleExfil = z(Ne); a.qEleRecharge = z(Ne); a.qEs = z(Ne); a.qEu = z(Ne); a.qEg = z(Ne);
        a.qTu = z(Ne); a.qTg = z(Ne); a.qEleTrans = z(Ne); a.qEleEvapo = z(Ne); a.qEleETA = z(Ne); a.iBeta = z(Ne);
        a.QeleSurf = z(3 * Ne); a.QeleSub = z(3 * Ne); a.QeleSurfTot = z(Ne); a.QeleSubTot = z(Ne); a.Qe2r_Surf = z(Ne);
        a.Qe2r_Sub = z(Ne); a.QrivSurf = z(Nr); a.QrivSub = z(Nr); a.QrivUp = z(Nr); a.QrivDown = z(Nr);
        a.y2LakeArea = z(Nl); a.QLakeSurf = z(Nl); a.QLakeSub = z(Nl); a.QLakeRivIn = z(Nl); a.QLakeRivOut = z(Nl);
        a.qLakeEvap = z(Nl); a.qLakePrcp = z(Nl);
        c->acc_effKH = z(Ne); c->acc_satn = z(Ne); c->acc_QsegSurf = z(Ns); c->acc_QsegSub = z(Ns);
        c->acc_alloc = true;
    }
    std::vector<const double *> src; std::vector<double *> dst; std::vector<size_t> len;
    acc_list(c, src, dst, len);
    for (size_t k = 0; k < src.size(); k++)
        if (len[k]) k_accumulate<<<(unsigned)std::min<size_t>((len[k] + 255) / 256, 1184), 256, 0, c->stream>>>(dst[k], src[k], len[k]);
    CK(cudaGetLastError());
    c->num_update++;
    return SHUD_OK;
}

int shud_b200_output_flush(shud_ctx *c, double tau, const shud_diag *o, int32_t *num_update) {
    if (!c || !o || !c->acc_alloc || c->num_update <= 0) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const double s = tau / c->num_update;
    // scale in place into a scratch copy laid out like the diag arrays, reset the accumulators, download
    std::vector<const double *> src; std::vector<double *> dst; std::vector<size_t> len;
    acc_list(c, src, dst, len);
    if (c->flush_tmp.empty()) {  // scratch laid out like the accumulators, allocated once (freed with the context)
        c->flush_tmp.assign(dst.size(), nullptr);
        for (size_t k = 0; k < dst.size(); k++)
            if (len[k]) c->flush_tmp[k] = dev_alloc<double>(c, len[k]);
    }
    std::vector<double *> &tmp = c->flush_tmp;
    for (size_t k = 0; k < dst.size(); k++) {
        if (!len[k]) continue;
        k_scale_copy<<<(unsigned)std::min<size_t>((len[k] + 255) / 256, 1184), 256, 0, c->stream>>>(tmp[k], dst[k], s, len[k]);
    }
    CK(cudaGetLastError());
    DevDiag t{};
    double **slots[] = {&t.qEleInfil, &t.qEleExfil, &t.qEleRecharge, &t.qEs, &t.qEu, &t.qEg, &t.qTu, &t.qTg, &t.qEleTrans,
                        &t.qEleEvapo, &t.qEleETA, &t.iBeta, &t.QeleSurf, &t.QeleSub, &t.QeleSurfTot, &t.QeleSubTot,
                        &t.Qe2r_Surf, &t.Qe2r_Sub, &t.QrivSurf, &t.QrivSub, &t.QrivUp, &t.QrivDown, &t.y2LakeArea,
                        &t.QLakeSurf, &t.QLakeSub, &t.QLakeRivIn, &t.QLakeRivOut, &t.qLakeEvap, &t.qLakePrcp};
    for (size_t k = 0; k < 29; k++) *slots[k] = tmp[k];
    int rc = download_diag(c, t, tmp[29], tmp[30], tmp[31], tmp[32], o);
    if (rc) return rc;  // the accumulators and the update count are untouched: the flush can be repeated
    for (size_t k = 0; k < dst.size(); k++)
        if (len[k]) CK(cudaMemsetAsync(dst[k], 0, sizeof(double) * len[k], c->stream));  // reset (Model_Control.cpp:958-960)
    if (num_update) *num_update = c->num_update;
    c->num_update = 0;
    return SHUD_OK;
}

int shud_b200_get_diag(shud_ctx *c, const shud_diag *o) {
    if (!c || !o || !c->diag_alloc) return SHUD_ERR_ARG;
    return download_diag(c, c->diag, c->m.effKH, c->m.satn, c->m.QsegSurf, c->m.QsegSub, o);
}

static int download_diag(shud_ctx *c, const DevDiag &d, const double *effKH, const double *satn, const double *qsegS,
                         const double *qsegG, const shud_diag *o) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const int Ne = c->Ne, Nr = c->Nr, Ns = c->Ns, Nl = c->Nl;
    std::vector<double> h;
    auto cell = [&](const double *dsrc, double *dst, int nblk) -> int {
        if (!dst) return 0;
        h.resize((size_t)nblk * Ne);
        if (cudaMemcpy(h.data(), dsrc, sizeof(double) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
        for (int b = 0; b < nblk; b++)
            for (int i = 0; i < Ne; i++) dst[(size_t)b * Ne + c->cperm[i]] = h[(size_t)b * Ne + i];
        return 0;
    };
    auto riv = [&](const double *dsrc, double *dst) -> int {
        if (!dst || Nr == 0) return 0;
        h.resize(Nr);
        if (cudaMemcpy(h.data(), dsrc, sizeof(double) * Nr, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
        for (int i = 0; i < Nr; i++) dst[c->rperm[i]] = h[i];
        return 0;
    };
    auto seg = [&](const double *dsrc, double *dst) -> int {
        if (!dst || Ns == 0) return 0;
        h.resize(Ns);
        if (cudaMemcpy(h.data(), dsrc, sizeof(double) * Ns, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
        for (int i = 0; i < Ns; i++) dst[c->sperm[i]] = h[i];
        return 0;
    };
    auto lake = [&](const double *dsrc, double *dst) -> int {
        if (!dst || Nl == 0) return 0;
        return cudaMemcpy(dst, dsrc, sizeof(double) * Nl, cudaMemcpyDeviceToHost) != cudaSuccess;
    };
    int bad = 0;
    bad |= cell(d.qEleInfil, o->qEleInfil, 1); bad |= cell(d.qEleExfil, o->qEleExfil, 1);
    bad |= cell(d.qEleRecharge, o->qEleRecharge, 1); bad |= cell(d.qEs, o->qEs, 1); bad |= cell(d.qEu, o->qEu, 1);
    bad |= cell(d.qEg, o->qEg, 1); bad |= cell(d.qTu, o->qTu, 1); bad |= cell(d.qTg, o->qTg, 1);
    bad |= cell(d.qEleTrans, o->qEleTrans, 1); bad |= cell(d.qEleEvapo, o->qEleEvapo, 1);
    bad |= cell(d.qEleETA, o->qEleETA, 1); bad |= cell(d.iBeta, o->iBeta, 1);
    bad |= cell(effKH, o->u_effKH, 1); bad |= cell(satn, o->u_satn, 1);
    bad |= cell(d.QeleSurf, o->QeleSurf, 3); bad |= cell(d.QeleSub, o->QeleSub, 3);
    bad |= cell(d.QeleSurfTot, o->QeleSurfTot, 1); bad |= cell(d.QeleSubTot, o->QeleSubTot, 1);
    bad |= cell(d.Qe2r_Surf, o->Qe2r_Surf, 1); bad |= cell(d.Qe2r_Sub, o->Qe2r_Sub, 1);
    bad |= seg(qsegS, o->QsegSurf); bad |= seg(qsegG, o->QsegSub);
    bad |= riv(d.QrivSurf, o->QrivSurf); bad |= riv(d.QrivSub, o->QrivSub); bad |= riv(d.QrivUp, o->QrivUp);
    bad |= riv(d.QrivDown, o->QrivDown);
    bad |= lake(d.y2LakeArea, o->y2LakeArea); bad |= lake(d.QLakeSurf, o->QLakeSurf);
    bad |= lake(d.QLakeSub, o->QLakeSub); bad |= lake(d.QLakeRivIn, o->QLakeRivIn);
    bad |= lake(d.QLakeRivOut, o->QLakeRivOut); bad |= lake(d.qLakeEvap, o->qLakeEvap);
    bad |= lake(d.qLakePrcp, o->qLakePrcp);
    return bad ? SHUD_ERR_CUDA : SHUD_OK;
}

}  // extern "C"

#include "shud_land.cuh"
