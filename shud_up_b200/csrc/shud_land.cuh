// shud_land.cuh - the land-surface step on the device, included by shud_rhs.cu.
//
// Replaces the per-cell loops of Model_Data::updateforcing -> tReadForcing (src/ModelData/MD_ET.cpp:14-281) and
// Model_Data::ET (MD_ET.cpp:282-342), which the reference runs on the host once per ET step (src/Model/shud.cpp:
// 106-109), with one kernel that writes the RHS's forcing arrays in place: the 9*Ne doubles of
// shud_b200_set_forcing never cross PCIe, and the host keeps only O(stations + classes) work per step (time-series
// lookups, solarPosition()).  One thread per cell, ~200 B/cell of HBM traffic, 2 exp + 4 log per cell: HBM-bound.
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

__global__ void __launch_bounds__(256) k_land(DevMesh m, DevLand L, int tsr_n, double tsr_den, double DT_min, CryoStep cs) {
    __shared__ double s_sx[kTsrSmem], s_sy[kTsrSmem], s_sz[kTsrSmem], s_wdt[kTsrSmem];
    const double *t_forc = L.tab, *t_lai = t_forc + 5 * L.nforc, *t_mf = t_lai + L.nlc;
    const double *g_sx = t_mf + L.nmf, *g_sy = g_sx + L.tsr_cap, *g_sz = g_sy + L.tsr_cap, *g_wdt = g_sz + L.tsr_cap;
    const int ns = tsr_n < kTsrSmem ? tsr_n : kTsrSmem;
    for (int k = threadIdx.x; k < ns; k += blockDim.x) {
        s_sx[k] = g_sx[k]; s_sy[k] = g_sy[k]; s_sz[k] = g_sz[k]; s_wdt[k] = g_wdt[k];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.Ne) return;
    // ---------------- tReadForcing, MD_ET.cpp:21-281 ----------------
    const int idx = L.iForc[i] - 1;
    const double *row = t_forc + 5 * idx;
    double t_prcp = row[0] * L.cPrep;
    const double t0 = row[1], Zt = L.forc_z[idx], Zi = m.z_surf[i];
    double t_temp;  // TemperatureOnElevation, Equations.hpp:66-72
    if (fabs(Zi - kL_NA) < kZERO || fabs(Zt - kL_NA) < kZERO) t_temp = t0;
    else t_temp = t0 + (Zt - Zi) * kL_dTdZ;
    t_temp = t_temp + L.cTemp;
    const double lai = t_lai[L.iLC[i] - 1] * L.cLAItsd;
    const double mf = t_mf[L.iMF[i] - 1] * L.cMF / 1440.;
    const double dswrf_h = row[4];
    double dswrf_t = dswrf_h, factor = 1.0;
    if (L.tsr) {
        if (tsr_n < 0) {
            factor = 0.0;
        } else {
            double num = 0.0;
            if (tsr_den > 0.0 && tsr_n > 0) {
                const double nx = L.nx[i], ny = L.ny[i], nz = L.nz[i];
                for (int k = 0; k < tsr_n; k++) {
                    const bool in_s = k < kTsrSmem;
                    const double wdt = in_s ? s_wdt[k] : g_wdt[k];
                    if (!(wdt > 0.0)) continue;
                    const double sx = in_s ? s_sx[k] : g_sx[k], sy = in_s ? s_sy[k] : g_sy[k], sz = in_s ? s_sz[k] : g_sz[k];
                    const double cosi = nx * sx + ny * sy + nz * sz;
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
                feff = num / tsr_den;
                if (!isfinite(feff) || !(feff > 0.0)) feff = 0.0;
                if (feff > L.cap) feff = L.cap;
            }
            factor = feff;
        }
        dswrf_t = dswrf_h * factor;
    }
    double t_rn = L.net ? dswrf_t : dswrf_t * (1 - L.albedo[i]);
    const double Uz = fabs(row[3]) + 0.001;
    double t_rh = row[2];
    t_prcp = t_prcp * 0.001 / 1440.;
    t_rn = t_rn * 1.0e-6;
    t_rh = l_min(l_max(t_rh, kL_ConstRH), 1.0);
    const double fixP = L.fixP[i];
    const double lambda = 2.501 - 0.002361 * t_temp;                    // LatentHeat
    const double Gamma = 0.0016286 * fixP / lambda;                     // PsychrometricConstant
    const double es = 0.6108 * exp(17.27 * t_temp / (t_temp + 237.3));  // VaporPressure_Sat
    const double ea = es * t_rh;
    const double ed = es - ea;
    const double tt = t_temp + 237.3;
    const double Delta = 4098. * es / (tt * tt);                        // SlopeSatVaporPressure
    const double rho = 3.486 * fixP / (275. + t_temp);                  // AirDensity
    const bool lake = (m.flags[i] & F_LAKE) != 0;
    double G;
    if (lake) G = 0.;
    else if (lai > 0) G = 0.4 * exp(-0.5 * lai) * t_rn;
    else G = 0.1 * t_rn;
    const double RG = t_rn - G;
    // WindProfile(2.0, Uz, windH, 0., ROUGHNESS_WATER)
    const double U2 = Uz * log((2.0 - 0.) / kL_RoughWater) / log((L.windH[i] - 0.) / kL_RoughWater);
    double pm_ow;  // PET_PM_openwater, is_sm_et.cpp:57-64
    {
        double ETp = (Delta * RG * kL_SecADay + Gamma * 6.43 * (1.0 + 0.536 * U2) * ed) / (Delta + Gamma);
        ETp = ETp / lambda;
        ETp = ETp * 0.001 / kL_SecADay;
        pm_ow = ETp;
    }
    const double qPotEvap = L.cETP * pm_ow * 60.;
    const double vgFrac = m.vegFrac[i];
    double qPotTran, etp;
    int err = 0;
    if (lake || lai <= 0.) {
        qPotTran = L.cETP * 0.;
        etp = qPotEvap;
    } else {
        const double hc = lai * 0.5, Zm = hc * 1.3333;
        const double d = 0.67 * hc, Z_om = 0.123 * hc, Z_ov = 0.0123 * hc;  // AerodynamicResistance(Uz, hc, Zm, Zm)
        const double ra = log(fabs(Zm - d) / Z_om) * log(fabs(Zm - d) / (Z_ov)) / (kL_Karman * kL_Karman * Uz);
        if (ra <= 0.0 || isnan(ra) || isinf(ra) || fabs(ra - kL_NA) < kZERO) err = 10;  // CheckNonZero -> myexit(ERRNAN)
        const double rs = 200. / lai;  // BulkSurfaceResistance(lai)
        const double E_rad = Delta * RG, E_air = rho * kL_Cp * ed / ra, r_sa = rs / ra;  // PET_Penman_Monteith
        double ETp = (E_rad + E_air) / (Delta + Gamma * (1 + r_sa));
        ETp = ETp / lambda;
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
            const double x = tacc / cs.nday;  // mean of the day that just ended
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
        fu_Sub = 1. - frozen_fraction(ab / cs.size_b, L.sub_max, L.sub_min);
        fu_Surf = 1. - frozen_fraction(as / cs.size_s, L.surf_max, L.surf_min);
    }
    double snStg = L.snow[i];
    const double snFrac = frozen_fraction(T, kL_Train, kL_Tsnow);
    const double snAcc = snFrac * prcp;
    double snMelt = (T > kL_To ? (T - kL_To) * mf : 0.);
    snMelt = l_min(l_max(0., snStg / DT_min), l_max(0., snMelt));
    snStg += (snAcc - snMelt) * DT_min;
    double icStg = (vgFrac > kZERO) ? (L.ics[i] / vgFrac) : 0.0;
    double icAcc, icEvap;
    if (lai > kZERO) {
        const double icMax = L.cISmax * kL_IcMax * lai;
        icAcc = l_min(prcp - snAcc, l_max(0., (icMax - icStg) / DT_min));
        icEvap = l_min(l_max(0., icStg / DT_min), qPotEvap);
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
    CK(cudaMallocHost((void **)&c->land_stage, sizeof(double) * ntab));
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
    CK(cudaStreamSynchronize(c->stream));  // the staging buffer is reused
    double *h = c->land_stage;
    size_t o = 0;
    memcpy(h + o, S->forc, sizeof(double) * 5 * d.nforc); o += 5 * (size_t)d.nforc;
    memcpy(h + o, S->lai, sizeof(double) * d.nlc); o += d.nlc;
    memcpy(h + o, S->mf, sizeof(double) * d.nmf); o += d.nmf;
    const double *smp[4] = {S->tsr_sx, S->tsr_sy, S->tsr_sz, S->tsr_wdt};
    for (int a = 0; a < 4; a++) {
        if (n > 0) memcpy(h + o, smp[a], sizeof(double) * n);
        o += d.tsr_cap;
    }
    CK(cudaMemcpyAsync(d.tab, h, sizeof(double) * o, cudaMemcpyHostToDevice, c->stream));
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
    k_land<<<(c->Ne + 255) / 256, 256, 0, c->stream>>>(c->m, d, S->tsr_n, S->tsr_den, S->dt_min, cs);
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
