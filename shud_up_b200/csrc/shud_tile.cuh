// shud_tile.cuh - the cell kernel of the SHUD right-hand side and its river tail (included by shud_rhs.cu).
//
//   k_tile: one block = one tile of 128 consecutive cells (consecutive along the Hilbert curve: a compact patch of
//   the mesh) = 4 lateral + 4 vertical warps, 4 blocks per SM.  Every per-cell input of the tile comes into shared
//   memory with SIX TMA operations issued by one thread: three 2-D tensor copies (cp.async.bulk.tensor, SASS
//   UTMALDG) - the 32 static double slices, the 8 forcing / carried-state slices, the 6 int slices, each family
//   laid out in global memory as one [slice][cell] array so that a tile is a box of it - and three bulk copies
//   (UBLKCP) of the Ysurf / Yunsat / Ygw slices of the caller's vector, all counted on one mbarrier.  No thread
//   issues a per-cell load for a staged input.  Lateral role: effKH, 3 overland + 3 groundwater edge fluxes, the
//   tile's river segments; vertical role: updateElement, infiltration, recharge, ET partition, balance equations.
//   effKH (src/Equations/Equations.cpp:116-134) is evaluated where it is used: by the lateral role for its own cell
//   (staged parameters), and again for the ~15 % of neighbours outside the tile from one packed 64-byte record per
//   cell (z_surf, z_bottom and the 5 parameters): no pre-pass (it was one launch and 60 B/cell of traffic).
//   The state-only river work rides in the vertical warps' wait for the hand-back: every block evaluates the Manning
//   flux of its share of the reaches (Flux_RiverDown, src/ModelData/MD_RiverFlux.cpp:5-63), and the lakes are dealt
//   to the first blocks (bank-edge sums, inflow, bathymetry, lake equation).
//   k_river_tail (programmatic dependent launch behind k_tile): per reach the sums of its segments' fluxes and of
//   its upstream reaches' flux and the stage equation (PassValue / f_applyDY, MD_f.cpp:157-179,228-240).
// No atomics on any flux; every sum in a fixed order; ydot is bit-reproducible run to run and bit-identical
// between a partition and the whole domain.
#pragma once
#ifdef RK_DEBUG
#include <cassert>
#define RK_ASSERT(c) assert(c)
#else
#define RK_ASSERT(c)
#endif

#ifndef RK_KH_PREPASS
#define RK_KH_PREPASS 0
#endif

namespace rk {

// suspend-time hint of mbarrier.try_wait [ns]: a waiting thread sleeps in hardware until the phase completes (or this
// long) instead of spinning through the issue slots of the warps that have work
#ifndef RK_WAIT_HINT
#define RK_WAIT_HINT 20000u
#endif

// double slices of a stage, in three families that are each ONE array [slice][ld] in global memory (a tile = a box):
//   forcing / carried state (8), statics (32); then the 3 slices of the caller's state vector and the computed effKH
enum {
    S_NETP = 0, S_PE, S_PT, S_LAI, S_FUS, S_FUB, S_SATN, S_EIC,                     // dyn family (8)
    S_AQD, S_SY, S_INFD, S_INFK, S_MACKV, S_HAF, S_THS, S_THR, S_THFC, S_BETA, S_KSV, S_VEG, S_IMP, S_WET, S_ROOT,
    S_ZS, S_ZB, S_DEP, S_AREA, S_E0, S_E1, S_E2, S_D0, S_D1, S_D2, S_R0, S_R1, S_R2,
    S_MACD, S_MACKH, S_VAF, S_KSH,                                                  // static family (32)
    S_YSF, S_YUS, S_YGW,                                                            // the state vector (3 bulk copies)
    S_KH,                                                                           // computed in place
    S_ND
};
constexpr int N_DYN = 8, N_STAT = 32, S_STAT0 = S_AQD;
static_assert(S_AQD == N_DYN && S_YSF == N_DYN + N_STAT, "slice families");
enum { I_NB0 = 0, I_NB1, I_NB2, I_FL, I_SEG0, I_SEGN, I_NI };  // int family (6): SEGN[i] = cell_seg_first[i + 1]

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
    const unsigned a = smem_u32(b);
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "MBAR_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        " @p bra MBAR_DONE;\n"
        " bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(a), "r"(parity), "r"(RK_WAIT_HINT) : "memory");
}
// the same on 32-bit shared-window addresses (nothing 64-bit to keep alive across a tile)
__device__ __forceinline__ void mbar_arrive_u(unsigned a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait_u(unsigned a, unsigned parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "MBAR_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        " @p bra MBAR_DONE;\n"
        " bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(a), "r"(parity), "r"(RK_WAIT_HINT) : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA 2-D tensor copy global -> shared: the box at column c0, row c1 of the tensor map (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// TMA prefetch of a box / of a contiguous range into L2 (no destination)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch(const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
// barrier among the RT threads of a team's lateral role (one warp: __syncwarp)
template <int RT>
__device__ __forceinline__ void lat_sync(int id) {
    if (RT == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(RT) : "memory");
}

// one packed record per cell: what an edge flux needs from a neighbour OUTSIDE the tile besides its state
struct __align__(64) NbRec {
    double zs, zb, aqd, macD, kmac, af, kmx, pad;  // a lake cell carries macD = 0: effKH = KsatH (Element.cpp:336-337)
};

}  // namespace rk

// Manning flux of reach r towards its downstream end, everything gathered from global memory
__device__ __forceinline__ double reach_down_flux(const DevMesh &m, const double *__restrict__ Yr, int r, int *err) {
    const double yraw = Yr[r];
    const double ystg = (m.r_bc[r] > 0) ? m.r_yBC[r] : yraw;
    const int down = m.r_down[r];
    double y_dn = 0., depth_dn = 0., slope_dn = 0.;
    if (down >= 0) {
        y_dn = (m.r_bc[down] > 0) ? m.r_yBC[down] : Yr[down];
        depth_dn = m.r_depth[down];
        slope_dn = m.r_slope[down];
    }
    // r_down device coding: >=0 downstream reach (device id); <0 the reference's outlet code
    return river_down(yraw, ystg, m.r_w0[r], m.r_bank[r], m.r_len[r], m.r_slope[r], m.r_depth[r], m.r_rough[r],
                      m.r_dist[r], down >= 0 ? 1 : down, m.r_toLake[r], y_dn, depth_dn, slope_dn, err);
}

// fixed-shape sum over 128 threads (4 whole warps, local index t) of a block (deterministic): warp shuffle tree, then warp 0 over
// the 4 warp partials.  `bar` = named barrier the 128 threads share.
__device__ __forceinline__ double sum128(double v, double *sm, int bar, int t) {
    asm volatile("bar.sync %0, 128;" ::"r"(bar) : "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((t & 31) == 0) sm[t >> 5] = v;
    asm volatile("bar.sync %0, 128;" ::"r"(bar) : "memory");
    double w = 0.;
    if (t < 32) {
        w = (t < 4) ? sm[t] : 0.;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    }
    return w;  // valid in thread t = 0
}

// one lake, by 128 threads (local index t) of a block (MD_f.cpp:16-17,44-47,180-191; bank edges MD_ElementFlux.cpp:46-53,
// 107-121; inflow MD_RiverFlux.cpp:24).  State only: needs nothing of the cell tiles.
template <bool DIAG>
__device__ __forceinline__ void lake_equation(const DevMesh &m, const DevDiag &d, const double *__restrict__ Y,
                                              double *__restrict__ DY, int l, double *sm, int t) {
    const size_t NE = (size_t)m.Ne, LD = (size_t)m.ld;
    const double *Yr = Y + 3 * NE;
    const double yl = Y[3 * NE + m.Nr + l];
    double qs = 0., qg = 0., qin = 0.;
    for (int k = m.l_bank_ptr[l] + t; k < m.l_bank_ptr[l + 1]; k += 128) {
        const int i = m.bank_cell[k], j = m.bank_j[k];
        const unsigned fl = m.flags[i];
        const double ysf = Y[i];
        const double isf = ysf < 0. ? 0. : ysf, zs = m.z_surf[i];
        const double ygw = (fl & F_HEADBC) ? m.ele_yBC[i] : Y[2 * NE + i];
        const double B = m.edge[j * LD + i];
        int e = 0;
        const double kh = eff_kh(ygw, m.aqd[i], m.macD[i], m.macKsatH[i], m.vAreaF[i], m.ksatH[i], &e);
        qs += weir_jtoi(m.l_zmin[l], yl < 0. ? 0. : yl, zs, isf, zs, 0.6, B, 0.01);
        // QLakeSub takes Q before the fu_Sub factor (MD_ElementFlux.cpp:121 precedes :153)
        qg += edge_sub(ygw, m.z_bottom[i], yl, m.l_yi0[l], kh, m.bank_kh[k], m.dist[j * LD + i], B);
    }
    for (int k = m.l_rin_ptr[l] + t; k < m.l_rin_ptr[l + 1]; k += 128) {
        int e2 = 0;
        qin += reach_down_flux(m, Yr, m.l_rin_idx[k], &e2);
    }
    qs = sum128(qs, sm, 15, t);
    qg = sum128(qg, sm, 15, t);
    qin = sum128(qin, sm, 15, t);
    if (t == 0) {
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

// ---------------------------------------------------------------------------------------------
// The two roles of a team on one staged tile (sd / si = the stage's double / int slices).  `sy` provides the
// team's synchronisation: lat_sync (lateral warps among themselves), over_arrive / over_wait (vertical ->
// lateral hand-over), back_arrive / back_wait (lateral -> vertical hand-back).
// ---------------------------------------------------------------------------------------------
#define SD(a) sd[(a) * RT + lc]
#define TILE_LOCALS()                                                                                          \
    const int Ne = m.Ne;                                                                                       \
    const size_t NE = (size_t)Ne, LD = (size_t)m.ld;                                                           \
    const int i0 = tile * RT;                                                                                  \
    const int i = i0 + lc;                                                                                     \
    const bool valid = i < Ne;                                                                                 \
    const int ic = valid ? i : Ne - 1; /* clamped index: tail threads read something harmless */               \
    const bool ytile = ((Ne & 1) == 0) && (i0 + RT <= Ne);                                                     \
    (void)LD; (void)NE;
template <bool DIAG, int RT, class SY>
__device__ __forceinline__ void vertical_role(const DevMesh &m, const DevDiag &d, const double *__restrict__ Y,
                                              double *__restrict__ DY, double *sd, const int *si, int tile, int lc, SY &sy) {
    using namespace rk;
    TILE_LOCALS()
        // ------------------------------ vertical role ------------------------------
        // runs in steps, each fetching only its own inputs from the staged slices and parking what a later
        // step needs back in shared memory: nothing is held in a register across the pow() calls
#define fl ((unsigned)si[I_FL * RT + lc])  /* re-read where needed: not worth a register across the pow() calls */
        // (both roles store the same final values into the slices they share: a slot never holds anything else)
        if (!ytile) { SD(S_YSF) = Y[ic]; SD(S_YUS) = Y[NE + ic]; }
        if (fl & F_HEADBC) SD(S_YGW) = m.ele_yBC[ic];
        else if (!ytile) SD(S_YGW) = Y[2 * NE + ic];
#define VFENCE() asm volatile("" ::: "memory")
        // ---- step 1: updateElement (2 pow) ----
        SoilState st;
        if (fl & F_LAKE) { st.deficit = 0.; st.theta = 0.; st.satn = 1.; st.satKr = 0.; }
        else st = cell_soil_state(SD(S_AQD), SD(S_THS), SD(S_THR), SD(S_BETA), SD(S_YUS), SD(S_YGW));
        VFENCE();
        // ---- step 2: infiltration / exfiltration / recharge; hand-over ----
        {
            CellVert v;
            v.satn = 1.; v.infil = v.exfil = v.rech = 0.;
            const double ysf = SD(S_YSF), netPrep = SD(S_NETP);
            if (!(fl & F_LAKE)) {
                CellParams p;
                CellForc f;
                p.aqd = SD(S_AQD); p.infD = SD(S_INFD); p.infKsatV = SD(S_INFK); p.macKsatV = SD(S_MACKV);
                p.hAreaF = SD(S_HAF); p.thetaR = SD(S_THR); p.thetaFC = SD(S_THFC); p.ksatV = SD(S_KSV);
                f.netPrep = netPrep; f.fuSurf = SD(S_FUS); f.fuSub = SD(S_FUB);
                cell_soil_flux(p, f, ysf, SD(S_YUS), SD(S_YGW), st, v);
            }
            const double isf2 = ysf - v.infil + v.exfil;
            // values handed between the roles live in input slices this role has finished with
            SD(S_NETP) = netPrep - v.infil + v.exfil;   // P1
            SD(S_INFD) = v.rech - v.exfil;              // G1
            SD(S_INFK) = dmax(0., isf2);                // ponding left for the river weir
            sy.over_arrive();                        // hand-over: the lateral warps wait for it
            SD(S_MACKV) = v.infil - v.rech;             // first difference of ydot[unsat]
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
            const double potEvap = SD(S_PE);
            if (fl & F_LAKE) {
                v.Es = v.Eu = v.Eg = v.Tu = v.Tg = 0.; v.eic = 0.; v.iBeta = 0.;
            } else {
                CellParams p;
                CellForc f;
                p.thetaS = SD(S_THS); p.thetaR = SD(S_THR); p.vegFrac = SD(S_VEG); p.impAF = SD(S_IMP);
                p.wetland = SD(S_WET); p.rootReach = SD(S_ROOT);
                f.potEvap = potEvap; f.potTran = SD(S_PT); f.lai = SD(S_LAI);
                cell_et(p, f, SD(S_YSF), SD(S_YUS), SD(S_YGW), SD(S_SATN), SD(S_EIC), v);
            }
            if (valid) {
                if (v.err) raise_err(m.err, v.err, i + 1);
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
        //      P1 - SurfTot/area and G1 - SubTot/area in the P1 / G1 slices (hand-back barrier). ----
        // the five ET terms wait in input slices this role has finished with (nothing is held in a register across
        // the idle work and the barrier)
        SD(S_PE) = v.Es; SD(S_PT) = v.Eg; SD(S_LAI) = v.Tg; SD(S_VEG) = v.Eu; SD(S_IMP) = v.Tu;
        sy.idle_work(lc);  // whatever else the block has for warps that would only wait here (state-only river work)
        sy.back_wait();
        if (valid) {
            const double area = SD(S_AREA), syield = SD(S_SY);
            double dsf = SD(S_NETP) - SD(S_PE);
            double dgw = SD(S_INFD) - SD(S_PT) - SD(S_LAI);
            if (fl & F_HEADBC) dgw = 0;
            else if (fl & F_FLUXBC) dgw += SHUD_DIVS(m.ele_QBC[i], area);
            if (fl & F_SS_SURF) dsf += SHUD_DIVS(m.qss[i], area);
            else if (fl & F_SS_GW) dgw += SHUD_DIVS(m.qss[i], area);
            dgw = SHUD_DIVS(dgw, syield);
            double dus = SD(S_MACKV) - SD(S_VEG) - SD(S_IMP);
            dus = SHUD_DIVS(dus, syield);
            if (fl & F_LAKE) { dsf = 0.; dus = 0.; dgw = 0.; }
            DY[i] = dsf;
            DY[NE + i] = dus;
            DY[2 * NE + i] = dgw;
        }
#undef VFENCE
#undef fl
}

template <bool DIAG, int RT, class SY>
__device__ __forceinline__ void lateral_role(const DevMesh &m, const DevDiag &d, const double *__restrict__ Y,
                                             double *__restrict__ DY, double *sd, const int *si, int tile, int lc, SY &sy) {
    using namespace rk;
    TILE_LOCALS()
        // ------------------------------ lateral role ------------------------------
        const unsigned fl = (unsigned)si[I_FL * RT + lc];
        const int q0 = si[I_SEG0 * RT], nsq = si[I_SEGN * RT + RT - 1] - q0;  // the tile's segment slots
        const bool has_seg = lc < nsq;
        int s_riv = 0, s_bc = 0;
        RK_ASSERT(q0 >= 0 && nsq >= 0 && q0 + nsq <= m.Ns);
        if (has_seg) { s_riv = __ldg(m.cs_riv + q0 + lc); s_bc = __ldg(m.cs_bc + q0 + lc); }
        RK_ASSERT(s_riv >= 0 && s_riv < (m.Nr > 0 ? m.Nr : 1));
        // the lines the edge loop gathers for neighbours outside the tile (record, Ysurf, Ygw): requested into
        // L1 now, so that the three gathers are not three serial round trips later
#ifndef RK_L1PF
#define RK_L1PF 0
#endif
#pragma unroll
        for (int j = 0; j < 3 && RK_L1PF; j++) {
            const int kk = si[(I_NB0 + j) * RT + lc];
            if (kk >= 0 && kk < Ne && (unsigned)(kk - i0) >= (unsigned)RT) {
                prefetch_l1(static_cast<const NbRec *>(m.nbrec) + kk);
                prefetch_l1(Y + kk);
                prefetch_l1(Y + 2 * NE + kk);
            }
        }
        if (!ytile) SD(S_YSF) = Y[ic];
        if (fl & F_HEADBC) SD(S_YGW) = m.ele_yBC[ic];
        else if (!ytile) SD(S_YGW) = Y[2 * NE + ic];
        int err = 0;
        // effKH of the own cell (a lake cell: KsatH, _Element::updateLakeElement, Element.cpp:336-337)
        {
            double kh = SD(S_KSH);
            if (RK_KH_PREPASS) {
                kh = SD(S_KH);
            } else if (!(fl & F_LAKE)) {
                int e = 0;
                kh = eff_kh(SD(S_YGW), SD(S_AQD), SD(S_MACD), SD(S_MACKH), SD(S_VAF), kh, &e);
                if (e && valid) err = e;
            }
            SD(S_KH) = kh;
            if (DIAG && valid) m.effKH[i] = kh;
        }
        sy.lat_sync();  // the tile's neighbour table (Ysurf, Ygw, z_surf, z_bottom, effKH) is complete
        // stage of the reach of this lane's segment slot (stage BC applied): its reach id has landed by now (no warp
        // waits for it ahead of the barrier), and the gather is in flight during the edge loop
        double s_yr = 0.;
        if (has_seg) s_yr = (s_bc > 0) ? m.r_yBC[s_riv] : Y[3 * NE + s_riv];
        if (!(fl & F_LAKE)) {
            const double ysf = SD(S_YSF);
            const double isf = ysf < 0. ? 0. : ysf;
            // one copy of the edge code, three trips (unrolled, the kernel outgrows the instruction cache)
#pragma unroll 1
            for (int j = 0; j < 3; j++) {
                double qs = 0., qg = 0.;
                const int kk = si[(I_NB0 + j) * RT + lc];
                RK_ASSERT(kk < Ne + m.Nhalo && kk >= -2 - m.nbank);
                const double ygw = SD(S_YGW), zs = SD(S_ZS), zb = SD(S_ZB), kh = SD(S_KH);
                if (kk >= 0) {
                    double nsf, ygw_n, zs_n, zb_n, kh_n;
                    const unsigned r = (unsigned)(kk - i0);
                    if (r < (unsigned)RT && kk < Ne) {  // inside the tile: shared memory (in a ragged last
                                                         // tile a halo id Ne+h also falls into [i0, i0+RT))
                        nsf = sd[S_YSF * RT + r]; ygw_n = sd[S_YGW * RT + r]; zs_n = sd[S_ZS * RT + r];
                        zb_n = sd[S_ZB * RT + r]; kh_n = sd[S_KH * RT + r];
                    } else if (kk < Ne) {
                        const double2 *rec = reinterpret_cast<const double2 *>(static_cast<const NbRec *>(m.nbrec) + kk);
                        nsf = Y[kk]; ygw_n = Y[2 * NE + kk];
                        if (m.has_headbc && (m.flags[kk] & F_HEADBC)) ygw_n = m.ele_yBC[kk];
                        const double2 a = __ldg(rec);
                        zs_n = a.x; zb_n = a.y;
                        if (RK_KH_PREPASS) {
                            kh_n = m.effKH[kk];
                        } else {
                            const double2 b = __ldg(rec + 1), c2 = __ldg(rec + 2), d2 = __ldg(rec + 3);
                            int e = 0;
                            kh_n = eff_kh(ygw_n, b.x, b.y, c2.x, c2.y, d2.x, &e);
                        }
                    } else {  // halo cell of a partition: state from the last halo exchange
                        const int h = kk - Ne;
                        nsf = m.h_state[2 * h]; ygw_n = m.h_state[2 * h + 1]; zs_n = m.h_zs[h]; zb_n = m.h_zb[h];
                        int e = 0;
                        kh_n = eff_kh(ygw_n, m.h_aqd[h], m.h_macD[h], m.h_macKsatH[h], m.h_vAreaF[h], m.h_ksatH[h], &e);
                    }
                    nsf = nsf < 0. ? 0. : nsf;
                    const double Bj = SD(S_E0 + j), dj = SD(S_D0 + j);
                    qs = edge_surface(isf, zs, nsf, zs_n, SD(S_DEP), dj, Bj, SD(S_R0 + j));
                    qg = edge_sub(ygw, zb, ygw_n, zb_n, kh, kh_n, dj, Bj);
                } else if (kk <= -2) {
                    // bank of a lake: weir over the shore + Darcy against the lake stage
                    const int slot = -2 - kk, l = m.bank_lake[slot];
                    const double yl = Y[3 * NE + m.Nr + l];
                    const double nsf = yl < 0. ? 0. : yl;
                    const double Bj = SD(S_E0 + j);
                    qs = weir_jtoi(m.l_zmin[l], nsf, zs, isf, zs, 0.6, Bj, 0.01);
                    qg = edge_sub(ygw, zb, yl, m.l_yi0[l], kh, m.bank_kh[slot], SD(S_D0 + j), Bj);
                } else if (!m.close_boundary) {
                    // open boundary (MD_ElementFlux.cpp:81-92,139-151)
                    const double d2e = m.dist2edge[j * LD + ic], dep = SD(S_DEP);
                    if (isf > dep) {
                        const double sl = isf / d2e * 0.5;
                        if (sl > 0.) qs = sqrt(sl) * cbrt(isf * isf * isf * isf * isf) * SD(S_E0 + j) / m.rough[ic];
                    }
                    if (ygw > dep * 10.) {
                        const double grad = ygw / d2e * 0.5;
                        if (grad > 0.) qg = kh * grad;
                    }
                }
                SD(S_E0 + j) = qs;  // the edge's statics are spent: its slices take the two fluxes
                SD(S_D0 + j) = qg * SD(S_FUB);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 3; j++) { SD(S_E0 + j) = 0.; SD(S_D0 + j) = 0.; }
        }
        // ---- river segments of the whole tile, one lane per segment slot (fun_Seg_sub / fun_Seg_surface,
        //      MD_RiverFlux.cpp:100-126): groundwater exchange and every load before the hand-over, only
        //      the weir (needs the ponding left after infiltration) after it.  Segment fluxes of the tile
        //      land in the spent roughness slices R0 (groundwater) and R1 (surface). ----
        double *const sq_g = sd + S_R0 * RT, *const sq_s = sd + S_R1 * RT;
        int s_lc = 0, s_sgm = 0;
        double s_zr = 0., s_zbk = 0., s_cwr = 0., s_len = 0.;
        if (has_seg) {
            const int q = q0 + lc;
            s_lc = __ldg(m.cs_cell + q) - i0; s_sgm = __ldg(m.cs_seg + q);
            RK_ASSERT(s_lc >= 0 && s_lc < RT && s_sgm >= 0 && s_sgm < m.Ns);
            s_zr = __ldg(m.cs_zr + q); s_zbk = __ldg(m.cs_zbk + q); s_cwr = __ldg(m.cs_cwr + q);
            s_len = __ldg(m.cs_len + q);
            const double qg = flux_r2e_gw(s_yr, s_zr, sd[S_YGW * RT + s_lc], sd[S_ZB * RT + s_lc],
                                          sd[S_KH * RT + s_lc], __ldg(m.cs_ksatH + q), s_len,
                                          __ldg(m.cs_bed + q)) * sd[S_FUB * RT + s_lc];
            m.QsegSub[s_sgm] = qg;
            sq_g[lc] = qg;
        }
        sy.over_wait();  // vertical role has handed over
        if (has_seg) {
            const double qs = weir_jtoi(sd[S_ZS * RT + s_lc], sd[S_INFK * RT + s_lc], s_zr, s_yr, s_zbk, s_cwr,
                                        s_len, sd[S_DEP * RT + s_lc]);
            m.QsegSurf[s_sgm] = qs;
            sq_s[lc] = qs;
        }
        for (int tq = RT + lc; tq < nsq; tq += RT) {  // tiles with more than RT segments (rare)
            const int q = q0 + tq;
            const int c2 = __ldg(m.cs_cell + q) - i0, sgm = __ldg(m.cs_seg + q), rr = __ldg(m.cs_riv + q);
            const double yr = (__ldg(m.cs_bc + q) > 0) ? m.r_yBC[rr] : Y[3 * NE + rr];
            const double zr = __ldg(m.cs_zr + q), len = __ldg(m.cs_len + q);
            const double qs = weir_jtoi(sd[S_ZS * RT + c2], sd[S_INFK * RT + c2], zr, yr, __ldg(m.cs_zbk + q),
                                        __ldg(m.cs_cwr + q), len, sd[S_DEP * RT + c2]);
            const double qg = flux_r2e_gw(yr, zr, sd[S_YGW * RT + c2], sd[S_ZB * RT + c2], sd[S_KH * RT + c2],
                                          __ldg(m.cs_ksatH + q), len, __ldg(m.cs_bed + q)) * sd[S_FUB * RT + c2];
            m.QsegSurf[sgm] = qs;
            m.QsegSub[sgm] = qg;
        }
        sy.lat_sync();  // segment fluxes of the tile are in shared memory
        {
            // element side of PassValue (MD_f.cpp:228-235): sum of this cell's segment fluxes, ascending id
            double e2rS = 0., e2rG = 0.;
            const int nseg = (int)(fl >> NSEG_SHIFT);
            if (nseg) {
                const int seg0 = si[I_SEG0 * RT + lc], tq0 = seg0 - q0;
                for (int kq = 0; kq < nseg; kq++) {
                    const int tq = tq0 + kq;
                    double qs, qg;
                    if (tq < RT) { qs = sq_s[tq]; qg = sq_g[tq]; }
                    else { const int sgm = __ldg(m.cs_seg + seg0 + kq); qs = m.QsegSurf[sgm]; qg = m.QsegSub[sgm]; }
                    e2rS += -qs;
                    e2rG += -qg;
                }
            }
            double surfTot = e2rS, subTot = e2rG;
            double Qs[3], Qg[3];
#pragma unroll
            for (int j = 0; j < 3; j++) { Qs[j] = SD(S_E0 + j); Qg[j] = SD(S_D0 + j); }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                surfTot += Qs[j];
                subTot += Qg[j];
                if (not_finite(Qs[j]) || not_finite(Qg[j])) err = err > 10 ? err : 10;
            }
            // first two terms of dYsf and dYgw; the vertical role subtracts its ET terms and finishes (hand-back)
            const double area = SD(S_AREA);
            SD(S_NETP) = SD(S_NETP) - SHUD_DIVS(surfTot, area);
            SD(S_INFD) = SD(S_INFD) - SHUD_DIVS(subTot, area);
            sy.back_arrive();
            if (valid) {
                if (err) raise_err(m.err, err, i + 1);
                if (DIAG) {
#pragma unroll
                    for (int j = 0; j < 3; j++) { d.QeleSurf[j * NE + i] = Qs[j]; d.QeleSub[j * NE + i] = Qg[j]; }
                    d.QeleSurfTot[i] = surfTot; d.QeleSubTot[i] = subTot; d.Qe2r_Surf[i] = e2rS; d.Qe2r_Sub[i] = e2rG;
                }
            }
        }
}
#undef SD
#undef TILE_LOCALS

// named-barrier synchronisation of the one team of a block: 1 lateral-internal, 2 hand-over, 3 hand-back
template <bool DIAG, int RT>
struct TeamBarriers {
    const DevMesh &m;
    const DevDiag &d;
    const double *Y;
    double *DY, *red;
    int share;  // reaches per block of the state-only river work (0: none in this launch)
    __device__ __forceinline__ void lat_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RT) : "memory"); }
    __device__ __forceinline__ void over_arrive() { asm volatile("bar.arrive 2, %0;" ::"n"(2 * RT) : "memory"); }
    __device__ __forceinline__ void over_wait() { asm volatile("bar.sync 2, %0;" ::"n"(2 * RT) : "memory"); }
    __device__ __forceinline__ void back_arrive() { asm volatile("bar.arrive 3, %0;" ::"n"(2 * RT) : "memory"); }
    __device__ __forceinline__ void back_wait() { asm volatile("bar.sync 3, %0;" ::"n"(2 * RT) : "memory"); }
    // The vertical warps reach the hand-back barrier ahead of the lateral warps: the block's share of the state-only
    // river work (Manning flux of its reaches; a lake, if the block has been dealt one) is done in that wait.
    __device__ __forceinline__ void idle_work(int t) {
#ifdef RK_NO_RIVER
        return;
#endif
        if (share <= 0) return;
        const int r = (int)blockIdx.x * share + t;
        if (t < share && r < m.Nr) {
            int err = 0;
            m.r_qdown[r] = reach_down_flux(m, Y + 3 * (size_t)m.Ne, r, &err);
            if (err) raise_err(m.err, err, r + 1);
        }
        for (int l = blockIdx.x; l < m.Nl; l += gridDim.x) lake_equation<DIAG>(m, d, Y, DY, l, red, t);
    }
};

// ---------------------------------------------------------------------------------------------
// The cell kernel: one block = one tile of RT = 128 cells = one team (4 lateral + 4 vertical warps), 4 blocks per SM.
// Thread 0 fetches the whole tile with six TMA operations (three 2-D tensor boxes of the [slice][cell] families,
// three bulk copies of the state-vector slices) onto one mbarrier.  The block's share of the state-only river work
// (Manning flux of ~Nr/ntiles reaches; a lake, if it has been dealt one) is done by the vertical warps while they
// would otherwise wait for the lateral warps' hand-back.
// ---------------------------------------------------------------------------------------------
namespace rk {
constexpr int RT = 128;
constexpr int STAGE_DBL = S_ND * RT, STAGE_INT = I_NI * RT;
constexpr int STAGE_BYTES = STAGE_DBL * 8 + STAGE_INT * 4;                 // 48,128
constexpr int LOAD_BYTES = (N_DYN + N_STAT + 3) * RT * 8 + I_NI * RT * 4;  // transaction bytes of a whole tile
constexpr int SMEM_BYTES = STAGE_BYTES + 64;                               // + mbarrier, lake-sum scratch
}  // namespace rk

template <bool DIAG>
__global__ void __launch_bounds__(2 * rk::RT, 4)
k_tile(const __grid_constant__ DevMesh m, const __grid_constant__ DevDiag d, const __grid_constant__ CUtensorMap map_dyn,
       const __grid_constant__ CUtensorMap map_stat, const __grid_constant__ CUtensorMap map_int,
       const double *__restrict__ Y, double *__restrict__ DY, int tile0, int river_share) {
    using namespace rk;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sd = reinterpret_cast<double *>(smem_raw);
    const int *si = reinterpret_cast<const int *>(sd + STAGE_DBL);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + STAGE_BYTES);
    double *red = reinterpret_cast<double *>(smem_raw + STAGE_BYTES + 16);
    const int tid = threadIdx.x;
    const int tile = tile0 + (int)blockIdx.x;
    const size_t NE = (size_t)m.Ne;
    // the river tail kernel may be scheduled as soon as every block of this grid is running
    asm volatile("griddepcontrol.launch_dependents;");
#ifndef RK_TMA
#define RK_TMA 1
#endif
#if RK_TMA
    if (tid == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const size_t i0 = (size_t)tile * RT;
        const bool ytile = ((m.Ne & 1) == 0) && (i0 + RT <= NE);
        mbar_expect_tx(full, (unsigned)(LOAD_BYTES - (ytile ? 0 : 3 * RT * 8) + (RK_KH_PREPASS ? RT * 8 : 0)));
        if (RK_KH_PREPASS) tma_load(smem_raw + S_KH * (RT * 8), m.effKH + i0, RT * 8, full);
        tma_load_2d(smem_raw + S_STAT0 * (RT * 8), &map_stat, (int)i0, 0, full);
        tma_load_2d(smem_raw, &map_dyn, (int)i0, 0, full);
        tma_load_2d(smem_raw + STAGE_DBL * 8, &map_int, (int)i0, 0, full);
        if (ytile) {
#pragma unroll
            for (int b = 0; b < 3; b++) tma_load(smem_raw + (S_YSF + b) * (RT * 8), Y + b * NE + i0, RT * 8, full);
        }
    }
    __syncthreads();  // the mbarrier is initialised for everybody
    mbar_wait(full, 0);
#else
    {
        // per-thread 8-byte asynchronous copies (LDGSTS): thread t brings cell (t & 127) of every other slice
        const size_t LD = (size_t)m.ld, i0 = (size_t)tile * RT;
        const int c = tid & (RT - 1), half = tid >> 7;
        const double *dynb = m.netPrep, *statb = m.aqd;
#pragma unroll
        for (int a = 0; a < (N_DYN + N_STAT) / 2; a++) {
            const int sl = 2 * a + half;
            const double *src = (sl < N_DYN ? dynb + (size_t)sl * LD : statb + (size_t)(sl - N_DYN) * LD) + i0 + c;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sd + sl * RT + c)), "l"(src) : "memory");
        }
        const int *intb = m.nbr;
#pragma unroll
        for (int a = 0; a < I_NI / 2; a++) {
            const int sl = 2 * a + half;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(si + sl * RT + c)), "l"(intb + (size_t)sl * LD + i0 + c) : "memory");
        }
        {
            const size_t ic = (i0 + c < NE) ? i0 + c : NE - 1;
            if (half == 0) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sd + S_YSF * RT + c)), "l"(Y + ic) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sd + S_YGW * RT + c)), "l"(Y + 2 * NE + ic) : "memory");
            } else {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sd + S_YUS * RT + c)), "l"(Y + NE + ic) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        (void)full;
    }
    __syncthreads();
#endif
    TeamBarriers<DIAG, RT> sy{m, d, Y, DY, red, river_share};
    if (tid >= RT) vertical_role<DIAG, RT>(m, d, Y, DY, sd, si, tile, tid - RT, sy);
    else lateral_role<DIAG, RT>(m, d, Y, DY, sd, si, tile, tid, sy);
}

// The reaches' stage equations behind the cell kernel (programmatic dependent launch: the index loads run under
// the cell kernel's last wave): per reach the sums of its segments' fluxes and of its upstream reaches' Manning
// flux, then f_applyDY's river part (PassValue MD_f.cpp:228-240, f_applyDY MD_f.cpp:157-179)
template <bool DIAG>
__global__ void __launch_bounds__(256) k_river_tail(const __grid_constant__ DevMesh m, const __grid_constant__ DevDiag d,
                                                    const double *__restrict__ Y, double *__restrict__ DY) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t NE = (size_t)m.Ne;
    if (r >= m.Nr) return;
    const int s0 = m.r_seg_ptr[r], s1 = m.r_seg_ptr[r + 1], u0 = m.r_up_ptr[r], u1 = m.r_up_ptr[r + 1];
    const int bc = m.r_bc[r];
    const double yraw = Y[3 * NE + r], w0 = m.r_w0[r], bank = m.r_bank[r], len = m.r_len[r];
    const double qbc = (bc < 0) ? m.r_qBC[r] : 0.;
    const RivGeom g = riv_geom(yraw, w0, bank);
    int up_first = (u0 < u1) ? m.r_up_idx[u0] : 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the cell kernel (segment fluxes, Manning fluxes) has completed
    const double qdown = m.r_qdown[r];
    double up = 0.;
    for (int k = u0; k < u1; k++) up += -m.r_qdown[k == u0 ? up_first : m.r_up_idx[k]];
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
}
