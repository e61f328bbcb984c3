// shud_land.cuh - the land-surface step on the device, included by shud_rhs.cu.
//
// Replaces the per-cell loops of Model_Data::updateforcing -> tReadForcing (src/ModelData/MD_ET.cpp:14-281) and
// Model_Data::ET (MD_ET.cpp:282-342), which the reference runs on the host once per ET step (src/Model/shud.cpp:
// 106-109), with one kernel that writes the RHS's forcing arrays in place: the 9*Ne doubles of
// shud_b200_set_forcing never cross PCIe, and the host keeps only O(stations + classes) work per step (time-series
// lookups, solarPosition()).  One thread per cell, ~200 B/cell of HBM traffic.  What depends only on the land-cover
// class (LAI, the two logarithms of the aerodynamic resistance, the canopy resistance, the soil-heat factor) is
// evaluated once per class by k_land_tables, same expressions and association, so a cell is left with 1 exp, 1 log
// and 8 divisions (the first version: 2 exp, 4 log, ~24 divisions = 1540 warp-instructions per 32 cells, issue-bound
// at 0.33 of the HBM roofline).  With SHUD_RCP (the build's default, shud_phys.cuh) divisors shared by several
// quotients or uniform over the cells are applied as reciprocals: <= 1.5 ulp per quotient, inside the 1e-12 tolerance.
// Helper formulas: src/Equations/is_sm_et.hpp / is_sm_et.cpp, Equations.hpp:66-72, functions.hpp:191-201; the
// constants are those of src/Model/Macros.hpp:43-83.  Parity: tests/test_land_gpu.py against sequences dumped
// from the reference itself (tests/golden/{ccw,qhh}.land.npz).
#pragma once

namespace {

constexpr double kL_SecADay = 86400, kL_dTdZ = 0.0065, kL_Tsnow = -3.0, kL_Train = 1.0, kL_To = 0.0;
constexpr double kL_RoughWater = 0.00137, kL_ConstRH = 0.01, kL_IcMax = 0.0002, kL_Karman = 0.4, kL_Cp = 1.013e-3;
constexpr double kL_NA = -9999;
constexpr int kTsrSmem = 64;

__device__ __forceinline__ double l_min(double a, double b) { return a > b ? b : a; }  // functions.hpp:117-123
__device__ __forceinline__ double l_max(double a, double b) { return a < b ? b : a; }
__device__ __forceinline__ double frozen_fraction(double T, double high, double low) {  // functions.hpp:191-201
    if (T > high) return 0;
    if (T < low) return 1;
    return l_min(1.0, l_max((high - T) / (high - low), 0.0));
}

#ifdef SHUD_RCP
#define L_RCP(x, d, rd) ((x) * (rd))  // x / d with the reciprocal of d at hand
#else
#define L_RCP(x, d, rd) ((x) / (d))
#endif

// The step's tables, once per step: pulled from the pinned staging buffer (mapped host memory: a few hundred doubles
// over PCIe, no copy-engine operation on the stream) into L.tab, and per land-cover class
// [lai | 0.4 exp(-lai/2) | log(.)log(.) of AerodynamicResistance | 200 / lai] into L.cls.  k_land is launched
// programmatically dependent: its per-cell loads run under this kernel, it waits before its first table read.
__global__ void k_land_tables(DevLand L, const double *__restrict__ stage, int ntab) {
    asm volatile("griddepcontrol.launch_dependents;");
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ntab; k += gridDim.x * blockDim.x) L.tab[k] = stage[k];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.nlc) return;
    const double lai = stage[5 * L.nforc + c] * L.cLAItsd;
    double gfac = 0.1, LL = 0., rs = 0.;
    if (lai > 0) {
        gfac = 0.4 * exp(-0.5 * lai);
        const double hc = lai * 0.5, Zm = hc * 1.3333;
        const double d = 0.67 * hc, Z_om = 0.123 * hc, Z_ov = 0.0123 * hc;  // AerodynamicResistance(Uz, hc, Zm, Zm)
        LL = log(fabs(Zm - d) / Z_om) * log(fabs(Zm - d) / (Z_ov));
        rs = 200. / lai;  // BulkSurfaceResistance(lai)
    }
    L.cls[c] = lai; L.cls[L.nlc + c] = gfac; L.cls[2 * L.nlc + c] = LL; L.cls[3 * L.nlc + c] = rs;
}

#ifndef LAND_MINB
#define LAND_MINB 8
#endif
#ifndef LAND_BLOCK
#define LAND_BLOCK 128
#endif
__global__ void __launch_bounds__(LAND_BLOCK, LAND_MINB) k_land(DevMesh m, DevLand L, int tsr_n, double tsr_den, double DT_min, CryoStep cs) {
    // sun samples of the step; a sample the reference skips for every cell (weight or cos(zenith) floor) gets weight 0
    __shared__ double s_sx[kTsrSmem], s_sy[kTsrSmem], s_sz[kTsrSmem], s_wdt[kTsrSmem], s_den[kTsrSmem], s_rden[kTsrSmem];
    const double *t_forc = L.tab, *t_lai = t_forc + 5 * L.nforc, *t_mf = t_lai + L.nlc;
    const double *g_sx = t_mf + L.nmf, *g_sy = g_sx + L.tsr_cap, *g_sz = g_sy + L.tsr_cap, *g_wdt = g_sz + L.tsr_cap;
    const int ns = tsr_n < kTsrSmem ? tsr_n : kTsrSmem;
    // the cell's own inputs first: their loads are in flight while the block stages the sun samples
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < m.Ne;
    const int ic = valid ? i : m.Ne - 1;
    const int idx = __ldg(L.iForc + ic) - 1, lc = __ldg(L.iLC + ic) - 1, imf = __ldg(L.iMF + ic) - 1;
    const unsigned cflags = __ldg(m.flags + ic);
    const double Zi = __ldg(m.z_surf + ic), fixP = __ldg(L.fixP + ic), windH = __ldg(L.windH + ic), vgFrac = __ldg(m.vegFrac + ic);
    const double albedo = L.net ? 0. : __ldg(L.albedo + ic);
    const double snow0 = L.snow[ic], ics0 = L.ics[ic];
    double nx = 0., ny = 0., nz = 0.;
    if (L.tsr && tsr_n > 0 && tsr_den > 0.0) { nx = __ldg(L.nx + ic); ny = __ldg(L.ny + ic); nz = __ldg(L.nz + ic); }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the step's tables are in place (k_land_tables)
    for (int k = threadIdx.x; k < ns; k += blockDim.x) {
        const double sz = g_sz[k];
        double denom = sz, wdt = g_wdt[k];
        if (denom < L.cosz_min) denom = L.cosz_min;
        if (!(wdt > 0.0) || !(denom > 0.0) || !isfinite(denom)) { wdt = 0.0; denom = 1.0; }
        s_sx[k] = g_sx[k]; s_sy[k] = g_sy[k]; s_sz[k] = sz; s_wdt[k] = wdt; s_den[k] = denom; s_rden[k] = 1.0 / denom;
    }
    __syncthreads();
    const double rDT = 1.0 / DT_min, r1440 = 1.0 / 1440., r_tsr_den = 1.0 / tsr_den;
    if (!valid) return;
    // ---------------- tReadForcing, MD_ET.cpp:21-281 ----------------
    const double *row = t_forc + 5 * idx;
    double t_prcp = row[0] * L.cPrep;
    const double t0 = row[1], Zt = L.forc_z[idx];
    double t_temp;  // TemperatureOnElevation, Equations.hpp:66-72
    if (fabs(Zi - kL_NA) < kZERO || fabs(Zt - kL_NA) < kZERO) t_temp = t0;
    else t_temp = t0 + (Zt - Zi) * kL_dTdZ;
    t_temp = t_temp + L.cTemp;
    const double lai = __ldg(L.cls + lc);  // t_lai[lc] * cLAItsd
    const double mf = L_RCP(t_mf[imf] * L.cMF, 1440., r1440);
    const double dswrf_h = row[4];
    double dswrf_t = dswrf_h, factor = 1.0;
    if (L.tsr) {
        if (tsr_n < 0) {
            factor = 0.0;
        } else {
            double num = 0.0;
            if (tsr_den > 0.0 && tsr_n > 0) {
                for (int k = 0; k < ns; k++) {
                    const double wdt = s_wdt[k];
                    if (!(wdt > 0.0)) continue;
                    const double cosi = nx * s_sx[k] + ny * s_sy[k] + nz * s_sz[k];
                    if (!(cosi > 0.0) || !isfinite(cosi)) continue;
                    double fk = L_RCP(cosi, s_den[k], s_rden[k]);
                    if (!isfinite(fk) || !(fk > 0.0)) continue;
                    if (fk > L.cap) fk = L.cap;
                    num += wdt * fk;
                }
                for (int k = kTsrSmem; k < tsr_n; k++) {  // more samples than the staged ones (rare): global tables
                    const double wdt = g_wdt[k];
                    if (!(wdt > 0.0)) continue;
                    const double sz = g_sz[k];
                    const double cosi = nx * g_sx[k] + ny * g_sy[k] + nz * sz;
                    if (!(cosi > 0.0) || !isfinite(cosi)) continue;
                    double denom = sz;
                    if (denom < L.cosz_min) denom = L.cosz_min;
                    if (!(denom > 0.0) || !isfinite(denom)) continue;
                    double fk = cosi / denom;
                    if (!isfinite(fk) || !(fk > 0.0)) continue;
                    if (fk > L.cap) fk = L.cap;
                    num += wdt * fk;
                }
            }
            double feff = 0.0;
            if (tsr_den > 0.0) {
                feff = L_RCP(num, tsr_den, r_tsr_den);
                if (!isfinite(feff) || !(feff > 0.0)) feff = 0.0;
                if (feff > L.cap) feff = L.cap;
            }
            factor = feff;
        }
        dswrf_t = dswrf_h * factor;
    }
    double t_rn = L.net ? dswrf_t : dswrf_t * (1 - albedo);
    const double Uz = fabs(row[3]) + 0.001;
    double t_rh = row[2];
    t_prcp = L_RCP(t_prcp * 0.001, 1440., r1440);
    t_rn = t_rn * 1.0e-6;
    t_rh = l_min(l_max(t_rh, kL_ConstRH), 1.0);
    const double lambda = 2.501 - 0.002361 * t_temp;                    // LatentHeat
    const double rlambda = 1.0 / lambda;
    const double Gamma = L_RCP(0.0016286 * fixP, lambda, rlambda);      // PsychrometricConstant
    const double tt = t_temp + 237.3;
#ifdef SHUD_RCP
    const double rtt = 1.0 / tt;
    const double es = 0.6108 * exp(17.27 * t_temp * rtt);               // VaporPressure_Sat
    const double Delta = 4098. * es * (rtt * rtt);                      // SlopeSatVaporPressure
#else
    const double es = 0.6108 * exp(17.27 * t_temp / (t_temp + 237.3));
    const double Delta = 4098. * es / (tt * tt);
#endif
    const double ea = es * t_rh;
    const double ed = es - ea;
    const double rho = 3.486 * fixP / (275. + t_temp);                  // AirDensity
    const bool lake = (cflags & F_LAKE) != 0;
    double G;
    if (lake) G = 0.;
    else G = __ldg(L.cls + L.nlc + lc) * t_rn;  // lai > 0: 0.4 exp(-lai/2) t_rn, else 0.1 t_rn
    const double RG = t_rn - G;
    // WindProfile(2.0, Uz, windH, 0., ROUGHNESS_WATER)
    const double U2 = Uz * log((2.0 - 0.) / kL_RoughWater) / log(L_RCP(windH - 0., kL_RoughWater, 1.0 / kL_RoughWater));
    double pm_ow;  // PET_PM_openwater, is_sm_et.cpp:57-64
    {
        double ETp = (Delta * RG * kL_SecADay + Gamma * 6.43 * (1.0 + 0.536 * U2) * ed) / (Delta + Gamma);
        ETp = L_RCP(ETp, lambda, rlambda);
        ETp = L_RCP(ETp * 0.001, kL_SecADay, 1.0 / kL_SecADay);
        pm_ow = ETp;
    }
    const double qPotEvap = L.cETP * pm_ow * 60.;
    double qPotTran, etp;
    int err = 0;
    if (lake || lai <= 0.) {
        qPotTran = L.cETP * 0.;
        etp = qPotEvap;
    } else {
        // AerodynamicResistance(Uz, hc, Zm, Zm) = log(.) log(.) / (k^2 Uz), the two logarithms from the class table
        const double LL = __ldg(L.cls + 2 * L.nlc + lc), rs = __ldg(L.cls + 3 * L.nlc + lc);
#ifdef SHUD_RCP
        // ra <= 0, NaN or Inf <=> the same of LL (k^2 Uz is positive and finite): CheckNonZero -> myexit(ERRNAN)
        if (!(LL > 0.0) || isinf(LL)) err = 10;
        const double rra = (kL_Karman * kL_Karman * Uz) / LL;
        const double E_rad = Delta * RG, E_air = rho * kL_Cp * ed * rra, r_sa = rs * rra;  // PET_Penman_Monteith
#else
        const double ra = LL / (kL_Karman * kL_Karman * Uz);
        if (ra <= 0.0 || isnan(ra) || isinf(ra) || fabs(ra - kL_NA) < kZERO) err = 10;
        const double E_rad = Delta * RG, E_air = rho * kL_Cp * ed / ra, r_sa = rs / ra;
#endif
        double ETp = (E_rad + E_air) / (Delta + Gamma * (1 + r_sa));
        ETp = L_RCP(ETp, lambda, rlambda);
        ETp = ETp * 0.001;
        qPotTran = L.cETP * ETp * 60.;
        etp = qPotTran * vgFrac + qPotEvap * (1. - vgFrac);
        if (isnan(qPotTran)) err = 10;
    }
    // ---------------- ET, MD_ET.cpp:282-342 ----------------
    const double T = t_temp, prcp = t_prcp;
    double fu_Surf = 1., fu_Sub = 1.;
    if (L.cryo) {  // MD_ET.cpp:301-307; _AccTemp::push / getACC, AccTemperature.hpp:28-60
        const size_t ld = (size_t)m.ld;
        double tacc = L.tacc[i] + T, as = L.acc_s[i], ab = L.acc_b[i];
        if (cs.do_push) {
            const double x = L_RCP(tacc, cs.nday, cs.r_nday);  // mean of the day that just ended
            as += x;
            if (cs.pop_s) as -= L.ring_s[cs.slot_s * ld + i];
            L.ring_s[cs.slot_s * ld + i] = x;
            ab += x;
            if (cs.pop_b) ab -= L.ring_b[cs.slot_b * ld + i];
            L.ring_b[cs.slot_b * ld + i] = x;
            tacc = 0.;
            L.acc_s[i] = as;
            L.acc_b[i] = ab;
        }
        L.tacc[i] = tacc;
        fu_Sub = 1. - frozen_fraction(L_RCP(ab, cs.size_b, cs.r_size_b), L.sub_max, L.sub_min);
        fu_Surf = 1. - frozen_fraction(L_RCP(as, cs.size_s, cs.r_size_s), L.surf_max, L.surf_min);
    }
    double snStg = snow0;
    const double snFrac = frozen_fraction(T, kL_Train, kL_Tsnow);
    const double snAcc = snFrac * prcp;
    double snMelt = (T > kL_To ? (T - kL_To) * mf : 0.);
    snMelt = l_min(l_max(0., L_RCP(snStg, DT_min, rDT)), l_max(0., snMelt));
    snStg += (snAcc - snMelt) * DT_min;
    double icStg = (vgFrac > kZERO) ? (ics0 / vgFrac) : 0.0;
    double icAcc, icEvap;
    if (lai > kZERO) {
        const double icMax = L.cISmax * kL_IcMax * lai;
        icAcc = l_min(prcp - snAcc, l_max(0., L_RCP(icMax - icStg, DT_min, rDT)));
        icEvap = l_min(l_max(0., L_RCP(icStg, DT_min, rDT)), qPotEvap);
    } else {
        icAcc = 0.;
        icEvap = 0.;
    }
    icStg += (icAcc - icEvap) * DT_min;
    L.ics[i] = icStg * vgFrac;
    L.snow[i] = snStg;
    // what shud_b200_set_forcing would have uploaded
    m.netPrep[i] = (1. - snFrac) * prcp + snMelt - icAcc * vgFrac;
    m.potEvap[i] = qPotEvap;
    m.potTran[i] = qPotTran;
    m.lai[i] = lai;
    m.fuSurf[i] = fu_Surf;
    m.fuSub[i] = fu_Sub;
    m.eic[i] = icEvap * vgFrac;
    // kept for the host (Print_Ctrl arrays, water-balance diagnostics) and the lake means
    L.prep[i] = t_prcp;
    L.etp[i] = etp;
    L.temp[i] = t_temp;
    L.tmf[i] = mf;
    L.factor[i] = factor;
    if (err) raise_err(m.err, err, i + 1);
}

// lake-cell means of qPotEvap / qElePrep, ascending reference cell order (MD_f.cpp:16-17): one thread per lake
__global__ void k_lake_means(DevMesh m, DevLand L) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= m.Nl) return;
    double ev = 0., pr = 0.;
    const double n = L.lk_rnele[l];
    for (int k = L.lk_ptr[l]; k < L.lk_ptr[l + 1]; k++) {
        const int i = L.lk_cell[k];
        ev += m.potEvap[i] / n;
        pr += L.prep[i] / n;
    }
    m.l_evap_raw[l] = ev;
    m.l_prcp[l] = pr;
}

const int *up_cell_i(shud_ctx *c, const int32_t *src) {
    std::vector<int> h(c->ld, 1);
    for (int i = 0; i < c->Ne; i++) h[i] = src[c->cperm[i]];
    return dev_upload(c, h);
}

}  // namespace

extern "C" {

int shud_b200_land_create(shud_ctx *c, const shud_land *L) {
    if (!c || !L || L->nforc <= 0 || L->nlc <= 0 || L->nmf <= 0) return SHUD_ERR_ARG;
    if (!L->iForc || !L->iLC || !L->iMF || !L->Albedo || !L->FixPressure || !L->windH || !L->forc_z) return SHUD_ERR_ARG;
    if (L->terrain_radiation && (!L->nx || !L->ny || !L->nz)) return SHUD_ERR_ARG;
    if (L->cryosphere && (L->FT_surf_day < 1. || L->FT_sub_day < 1. || L->FT_surf_day > 366. || L->FT_sub_day > 366.))
        return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < c->Ne; i++)
        if (L->iForc[i] < 1 || L->iForc[i] > L->nforc || L->iLC[i] < 1 || L->iLC[i] > L->nlc || L->iMF[i] < 1 ||
            L->iMF[i] > L->nmf)
            return SHUD_ERR_ARG;
    DevLand &d = c->land;
    d.nforc = L->nforc; d.nlc = L->nlc; d.nmf = L->nmf;
    d.iForc = up_cell_i(c, L->iForc); d.iLC = up_cell_i(c, L->iLC); d.iMF = up_cell_i(c, L->iMF);
    d.albedo = up_cell(c, L->Albedo); d.fixP = up_cell(c, L->FixPressure); d.windH = up_cell(c, L->windH);
    if (L->terrain_radiation) { d.nx = up_cell(c, L->nx); d.ny = up_cell(c, L->ny); d.nz = up_cell(c, L->nz); }
    d.forc_z = dev_upload(c, std::vector<double>(L->forc_z, L->forc_z + L->nforc));
    d.cPrep = L->cPrep; d.cTemp = L->cTemp; d.cLAItsd = L->cLAItsd; d.cMF = L->cMF; d.cETP = L->cETP; d.cISmax = L->cISmax;
    d.net = L->radiation_is_net; d.tsr = L->terrain_radiation; d.cap = L->rad_factor_cap; d.cosz_min = L->rad_cosz_min;
    const size_t ld = (size_t)c->ld;
    d.snow = dev_alloc<double>(c, ld); d.ics = dev_alloc<double>(c, ld);
    d.prep = dev_alloc<double>(c, ld); d.etp = dev_alloc<double>(c, ld); d.temp = dev_alloc<double>(c, ld);
    d.tmf = dev_alloc<double>(c, ld); d.factor = dev_alloc<double>(c, ld);
    for (double *p : {d.snow, d.ics, d.prep, d.etp, d.temp, d.tmf, d.factor}) {
        if (!p) return SHUD_ERR_CUDA;
        CK(cudaMemset(p, 0, sizeof(double) * ld));
    }
    d.cryo = L->cryosphere ? 1 : 0;
    if (d.cryo) {
        d.Ls = (int)L->FT_surf_day; d.Lb = (int)L->FT_sub_day;
        d.surf_max = L->FT_surf_max; d.surf_min = L->FT_surf_min; d.sub_max = L->FT_sub_max; d.sub_min = L->FT_sub_min;
        d.tacc = dev_alloc<double>(c, ld); d.acc_s = dev_alloc<double>(c, ld); d.acc_b = dev_alloc<double>(c, ld);
        d.ring_s = dev_alloc<double>(c, ld * d.Ls); d.ring_b = dev_alloc<double>(c, ld * d.Lb);
        if (!d.tacc || !d.acc_s || !d.acc_b || !d.ring_s || !d.ring_b) return SHUD_ERR_CUDA;
        CK(cudaMemset(d.tacc, 0, sizeof(double) * ld)); CK(cudaMemset(d.acc_s, 0, sizeof(double) * ld));
        CK(cudaMemset(d.acc_b, 0, sizeof(double) * ld));
        CK(cudaMemset(d.ring_s, 0, sizeof(double) * ld * d.Ls)); CK(cudaMemset(d.ring_b, 0, sizeof(double) * ld * d.Lb));
        c->cryo_tstart = -9999.; c->cryo_nday = 0.;
        c->cryo_size_s = c->cryo_head_s = c->cryo_size_b = c->cryo_head_b = 0;
    }
    d.tsr_cap = 256;
    const size_t ntab = 5 * (size_t)d.nforc + d.nlc + d.nmf + 4 * (size_t)d.tsr_cap;
    d.tab = dev_alloc<double>(c, ntab);
    d.cls = dev_alloc<double>(c, 4 * (size_t)d.nlc);
    if (!d.tab || !d.cls) return SHUD_ERR_CUDA;
    CK(cudaHostAlloc((void **)&c->land_stage, sizeof(double) * 2 * ntab, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c->land_stage_dev, c->land_stage, 0));
    c->land_ntab = ntab;
    for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&c->land_ev[k], cudaEventDisableTiming));
    // lake -> its cells, ascending reference id (c->lake_cells is ascending)
    std::vector<int> ptr(c->Nl + 1, 0), cell;
    std::vector<double> rn(std::max(c->Nl, 1), 1.0);
    if (c->Nl > 0) {
        std::vector<std::vector<int>> per(c->Nl);
        for (int o : c->lake_cells) per[c->lake_of_cell[o]].push_back(c->cinv[o]);
        for (int l = 0; l < c->Nl; l++) {
            ptr[l + 1] = ptr[l] + (int)per[l].size();
            cell.insert(cell.end(), per[l].begin(), per[l].end());
            rn[l] = (double)c->lake_nele[l];
        }
    }
    d.lk_ptr = dev_upload(c, ptr); d.lk_cell = dev_upload(c, cell); d.lk_rnele = dev_upload(c, rn);
    c->has_land = true;
    return SHUD_OK;
}

int shud_b200_land_set_state(shud_ctx *c, const double *yEleSnow, const double *yEleIS) {
    if (!c || !c->has_land || !yEleSnow || !yEleIS) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    int rc = upload_perm(c, c->land.snow, yEleSnow, c->cperm);
    if (rc) return rc;
    return upload_perm(c, c->land.ics, yEleIS, c->cperm);
}

int shud_b200_land_step(shud_ctx *c, const shud_land_step *S) {
    if (!c || !c->has_land || !S || !S->forc || !S->lai || !S->mf || !(S->dt_min > 0.)) return SHUD_ERR_ARG;
    DevLand &d = c->land;
    const int n = S->tsr_n > 0 ? S->tsr_n : 0;
    if (n > d.tsr_cap || (n > 0 && (!S->tsr_sx || !S->tsr_sy || !S->tsr_sz || !S->tsr_wdt))) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    const int half = c->land_flip;
    c->land_flip ^= 1;
    if (c->land_ev_set[half]) CK(cudaEventSynchronize(c->land_ev[half]));  // the step before last has read this half
    double *h = c->land_stage + half * c->land_ntab;
    size_t o = 0;
    memcpy(h + o, S->forc, sizeof(double) * 5 * d.nforc); o += 5 * (size_t)d.nforc;
    memcpy(h + o, S->lai, sizeof(double) * d.nlc); o += d.nlc;
    memcpy(h + o, S->mf, sizeof(double) * d.nmf); o += d.nmf;
    const double *smp[4] = {S->tsr_sx, S->tsr_sy, S->tsr_sz, S->tsr_wdt};
    for (int a = 0; a < 4; a++) {
        if (n > 0) memcpy(h + o, smp[a], sizeof(double) * n);
        o += d.tsr_cap;
    }
    const int ntab = (int)o;
    CryoStep cs = {};
    cs.size_s = cs.size_b = 1.;
    if (d.cryo) {
        // _AccTemp::push(x, tnow): N_of_day++, and once >= 1440 min have passed the day's mean enters the queue; a
        // full queue drops its oldest value (ring: the new value takes the slot of the one it displaces)
        c->cryo_nday += 1.;
        cs.nday = c->cryo_nday;
        cs.do_push = (S->t - c->cryo_tstart) >= 1440.;
        if (cs.do_push) {
            if (c->cryo_size_s == d.Ls) { cs.pop_s = 1; cs.slot_s = c->cryo_head_s; c->cryo_head_s = (c->cryo_head_s + 1) % d.Ls; }
            else { cs.slot_s = (c->cryo_head_s + c->cryo_size_s) % d.Ls; c->cryo_size_s++; }
            if (c->cryo_size_b == d.Lb) { cs.pop_b = 1; cs.slot_b = c->cryo_head_b; c->cryo_head_b = (c->cryo_head_b + 1) % d.Lb; }
            else { cs.slot_b = (c->cryo_head_b + c->cryo_size_b) % d.Lb; c->cryo_size_b++; }
            c->cryo_tstart = S->t;
            c->cryo_nday = 0.;
        }
        cs.size_s = (double)c->cryo_size_s; cs.size_b = (double)c->cryo_size_b;
    }
    cs.r_nday = 1. / cs.nday; cs.r_size_s = 1. / cs.size_s; cs.r_size_b = 1. / cs.size_b;
    k_land_tables<<<(std::max(ntab, d.nlc) + 255) / 256, 256, 0, c->stream>>>(d, c->land_stage_dev + half * c->land_ntab, ntab);

    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((c->Ne + LAND_BLOCK - 1) / LAND_BLOCK); cfg.blockDim = dim3(LAND_BLOCK); cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = c->use_pdl ? 1 : 0;
        if (cudaLaunchKernelEx(&cfg, k_land, c->m, d, S->tsr_n, S->tsr_den, S->dt_min, cs) != cudaSuccess) {
            cudaGetLastError();
            k_land<<<cfg.gridDim, cfg.blockDim, 0, c->stream>>>(c->m, d, S->tsr_n, S->tsr_den, S->dt_min, cs);
        }
    }
    CK(cudaEventRecord(c->land_ev[half], c->stream));  // (behind k_land: nothing between the two launches of the programmatic pair)
    c->land_ev_set[half] = 1;
    if (c->Nl > 0) k_lake_means<<<(c->Nl + 63) / 64, 64, 0, c->stream>>>(c->m, d);
    CK(cudaGetLastError());
    return SHUD_OK;
}

int shud_b200_land_get(shud_ctx *c, const shud_land_out *out) {
    if (!c || !c->has_land || !out) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const DevMesh &m = c->m;
    const DevLand &d = c->land;
    const double *src[] = {d.prep, m.potEvap, m.potTran, d.etp, m.lai, d.temp, d.tmf, m.netPrep, m.eic, m.fuSurf, m.fuSub,
                           d.factor, d.snow, d.ics};
    double *dst[] = {out->qElePrep, out->qPotEvap, out->qPotTran, out->qEleETP, out->t_lai, out->t_temp, out->t_mf,
                     out->qEleNetPrep, out->qEleE_IC, out->fu_Surf, out->fu_Sub, out->rn_factor, out->yEleSnow, out->yEleIS};
    std::vector<double> h(c->Ne);
    for (int a = 0; a < 14; a++) {
        if (!dst[a]) continue;
        CK(cudaMemcpy(h.data(), src[a], sizeof(double) * c->Ne, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c->Ne; i++) dst[a][c->cperm[i]] = h[i];
    }
    return SHUD_OK;
}

// checkpoint in the reference's initial-condition format, straight from the device (shud_io.cu has the formatter)
int shud_b200_write_ic(shud_ctx *c, const char *path, double t, const double *y_dev) {
    if (!c || !path || !y_dev) return SHUD_ERR_ARG;
    CK(cudaSetDevice(c->device));
    // the reference prints what summary() left in yEle*/yRivStg (shud.cpp:137-157): BC heads / stages, not the
    // solver's frozen rows; device order -> reference order on the device, one D2H
    std::vector<double> y(c->NY), is, sn;
    int rc = shud_b200_summary_dev(c, y_dev, y.data());
    if (rc) return rc;
    if (c->has_land) {
        is.resize(c->Ne); sn.resize(c->Ne);
        shud_land_out o = {};
        o.yEleIS = is.data(); o.yEleSnow = sn.data();
        rc = shud_b200_land_get(c, &o);
        if (rc) return rc;
    }
    return shud_b200_format_ic(path, t, c->Ne, c->Nr, c->Nl, c->has_land ? is.data() : nullptr,
                               c->has_land ? sn.data() : nullptr, y.data());
}

}  // extern "C"
