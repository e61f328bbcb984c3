/* TEST INFRASTRUCTURE ONLY — lives under oracle/, never linked into the product.
 *
 * Driver for the UNMODIFIED reference RHS.  oracle/Makefile compiles this file
 * together with the reference's own sources where they lie under /root/reference
 * (nothing is copied) into oracle/_ref/shud_ref_{serial,omp}.  It replays the
 * reference's set-up sequence (src/Model/shud.cpp:49-69 and :106-109), then calls
 * the reference f() (src/Model/f.cpp:2-32) and dumps, in one binary snapshot:
 *   - every static array the RHS reads, flattened AoS -> SoA  (the input of the
 *     C-ABI shud_b200_create(), include/shud_b200.h),
 *   - the forcing-step arrays, the carried state (u_satn, qEleE_IC),
 *   - y, and the reference's ydot and flux arrays for that y.
 * The snapshot is what pins oracle/shud_oracle.c and the CUDA path
 * (tests/golden/*.npz are made from it by tools/make_golden.py).
 *
 * f() is not a pure function of (t,y): the driver calls it
 * twice on the same y; the carried state is snapshotted between the calls and
 * the outputs are those of the second call.
 *
 * usage: shud_ref <prj> <out.bin> [--t MIN] [--state ic|rand:<seed>]
 *                 [--mutate a,b,..] [--time REPS] [--forcing-seq NSTEPS]
 *   --print-init FILE: write the state of this run with the reference's own checkpoint writer
 *   (Model_Data::PrintInit, src/ModelData/MD_update.cpp:268-299) after giving the canopy / snow buckets random
 *   values; the buckets are dumped as ic_yEleIS / ic_yEleSnow (pin of shud_b200_format_ic).
 *   --land-seq N [--land-t0 MIN] [--land-dt MIN] [--land-stride K] [--mutate cryo]: replay N consecutive land-surface steps (updateAllTimeSeries + updateforcing + ET) on a fresh
 *   model and dump, per step, everything the per-cell part consumes (station rows, LAI / melt-factor class
 *   values, the terrain-radiation solar samples of the forcing interval) and produces: the pin of the
 *   land-surface step.
 *   --forcing-seq N: additionally replay the reference's land-surface step for N consecutive ET steps of
 *   60 min (updateAllTimeSeries + updateforcing + ET, src/Model/shud.cpp:106-109) and dump what each hands to
 *   the RHS (fseq_<array>, [N][Ne]); these depend on the forcing files and the snow / interception buckets
 *   only, not on the solver state, so they can drive a full run without the reference at hand.
 *   run from a cwd that contains input/<prj>/ .
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <cmath>
#include <memory>
#include <fstream>
#include <sstream>
#include <iostream>
#include <algorithm>
#include <map>
/* The reference keeps a few RHS inputs private (t_lai, Kmax, u_satKr, the
 * time-series ring).  Access control does not change object layout, so the
 * driver (and only the driver) opens them up to read them out. */
#define private public
#include "Model_Data.hpp"
#undef private
#include "f.hpp"
#include "CommandIn.hpp"

/* globals the reference defines in src/Model/shud.cpp:19-30 (shud.cpp itself is
 * not compiled: it needs CVODE) */
double *uYsf;
double *uYus;
double *uYgw;
double *uYriv;
double *uYlake;
double *globalY;
double timeNow;
int dummy_mode = 0;
int global_fflush_mode = 0;
int global_implicit_mode = 1;
int global_verbose_mode = 1;
int lakeon = 0;

static FILE *g_out = nullptr;
static void put(const char *name, int dtype, int64_t n, const void *data) {
    char nm[48];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, sizeof nm - 1);
    fwrite(nm, 1, sizeof nm, g_out);
    int32_t dt = dtype;
    fwrite(&dt, 4, 1, g_out);
    fwrite(&n, 8, 1, g_out);
    fwrite(data, dtype == 0 ? 8 : 4, (size_t)n, g_out);
}
static void putd(const char *name, const std::vector<double> &v) { put(name, 0, (int64_t)v.size(), v.data()); }
static void puti(const char *name, const std::vector<int> &v) { put(name, 1, (int64_t)v.size(), v.data()); }
static void putd(const char *name, const double *p, int n) { put(name, 0, n, p); }
static void put1i(const char *name, int v) { put(name, 1, 1, &v); }
static void put1d(const char *name, double v) { put(name, 0, 1, &v); }

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static double urand() { /* splitmix64 -> [0,1) */
    uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

#define ELE_D(field)                                                         \
    {                                                                        \
        std::vector<double> v(Ne);                                           \
        for (int i = 0; i < Ne; i++) v[i] = MD->Ele[i].field;                \
        putd("ele_" #field, v);                                              \
    }
#define ELE_I(field)                                                         \
    {                                                                        \
        std::vector<int> v(Ne);                                              \
        for (int i = 0; i < Ne; i++) v[i] = MD->Ele[i].field;                \
        puti("ele_" #field, v);                                              \
    }
#define ELE_D3(field)                                                        \
    {                                                                        \
        std::vector<double> v(3 * (size_t)Ne);                               \
        for (int j = 0; j < 3; j++)                                          \
            for (int i = 0; i < Ne; i++) v[(size_t)j * Ne + i] = MD->Ele[i].field[j]; \
        putd("ele_" #field, v);                                              \
    }
#define ELE_I3(field)                                                        \
    {                                                                        \
        std::vector<int> v(3 * (size_t)Ne);                                  \
        for (int j = 0; j < 3; j++)                                          \
            for (int i = 0; i < Ne; i++) v[(size_t)j * Ne + i] = MD->Ele[i].field[j]; \
        puti("ele_" #field, v);                                              \
    }
#define RIV_D(field)                                                         \
    {                                                                        \
        std::vector<double> v(Nr);                                           \
        for (int i = 0; i < Nr; i++) v[i] = MD->Riv[i].field;                \
        putd("riv_" #field, v);                                              \
    }
#define RIV_I(field)                                                         \
    {                                                                        \
        std::vector<int> v(Nr);                                              \
        for (int i = 0; i < Nr; i++) v[i] = MD->Riv[i].field;                \
        puti("riv_" #field, v);                                              \
    }

static void fake_tsd(_TimeSeriesData &ts, int ncol_values) {
    /* a one-row in-memory series: getX(t,col) is ts[iNow][col]
     * (src/classes/TimeSeriesData.cpp:270-273) */
    ts.ts[0] = new double[ncol_values + 1];
    ts.ts[0][0] = 0.;
    for (int c = 1; c <= ncol_values; c++) ts.ts[0][c] = 0.;
    ts.iNow = 0;
    ts.iNext = 0;
}

static bool has(const std::string &list, const char *key) {
    std::stringstream ss(list);
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (tok == key) return true;
    return false;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <prj> <out.bin> [--t MIN] [--state ic|rand:<seed>] [--mutate a,b] [--time REPS]\n", argv[0]);
        return 2;
    }
    std::string prj = argv[1], outfn = argv[2], state = "ic", mutate = "", print_init = "";
    double t_arg = NAN, land_t0 = NAN, land_dt = 60.;
    int reps = 0, fseq = 0, lseq = 0, lstride = 1;
    for (int a = 3; a < argc; a++) {
        if (!strcmp(argv[a], "--t") && a + 1 < argc) t_arg = atof(argv[++a]);
        else if (!strcmp(argv[a], "--state") && a + 1 < argc) state = argv[++a];
        else if (!strcmp(argv[a], "--mutate") && a + 1 < argc) mutate = argv[++a];
        else if (!strcmp(argv[a], "--time") && a + 1 < argc) reps = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--forcing-seq") && a + 1 < argc) fseq = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--land-seq") && a + 1 < argc) lseq = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--land-dt") && a + 1 < argc) land_dt = atof(argv[++a]);
        else if (!strcmp(argv[a], "--print-init") && a + 1 < argc) print_init = argv[++a];
        else if (!strcmp(argv[a], "--land-t0") && a + 1 < argc) land_t0 = atof(argv[++a]);
        else if (!strcmp(argv[a], "--land-stride") && a + 1 < argc) lstride = atoi(argv[++a]);
        else { fprintf(stderr, "unknown arg %s\n", argv[a]); return 2; }
    }

    /* ---- src/Model/shud.cpp:359-364 and :49-69 ---- */
    CommandIn CLI;
    FileIn *fin = new FileIn;
    FileOut *fout = new FileOut;
    char a0[] = "shud";
    std::vector<char> a1(prj.begin(), prj.end());
    a1.push_back(0);
    char *av[] = {a0, a1.data(), nullptr};
    CLI.parse(2, av);
    CLI.setFileIO(fin, fout);
    Model_Data *MD = new Model_Data(fin, fout);
    MD->loadinput();
    MD->initialize();
    MD->CheckInputData();
    fout->updateFilePath();
    const int NY = MD->NumY;
    globalY = new double[NY];
    _N_VectorContent_Serial cu = {NY, 1, new double[NY]}, cdu = {NY, 1, new double[NY]};
    _generic_N_Vector vu = {&cu, nullptr, nullptr}, vdu = {&cdu, nullptr, nullptr};
    N_Vector udata = &vu, du = &vdu;
    MD->LoadIC();
    MD->SetIC2Y(udata);

    const int Ne = MD->NumEle, Nr = MD->NumRiv, Ns = MD->NumSegmt, Nl = MD->NumLake;
    double *Y = NV_DATA_S(udata), *DY = NV_DATA_S(du);

    /* ---- src/Model/shud.cpp:106-109 ---- */
    double t = std::isnan(t_arg) ? MD->CS.StartTime : t_arg;
    MD->updateAllTimeSeries(t);
    MD->updateforcing(t);
    MD->ET(t, t + 60.);

    /* ---- optional in-memory mutations: switch on branches no shipped basin
     * reaches (BC=0, SS=0, CloseBoundary=1, cryosphere=0) ---- */
    if (state.rfind("rand:", 0) == 0) g_rng ^= (uint64_t)atoll(state.c_str() + 5) * 0xD1342543DE82EF95ull;
    if (has(mutate, "openbnd")) MD->CS.CloseBoundary = 0;
    if (has(mutate, "frozen")) {
        for (int i = 0; i < Ne; i++) {
            MD->fu_Surf[i] = 0.1 + 0.9 * urand();
            MD->fu_Sub[i] = 0.1 + 0.9 * urand();
        }
    }
    if (has(mutate, "ss")) {
        for (int i = 0; i < Ne; i++) {
            if (i % 7 == 3) { MD->Ele[i].iSS = 1; MD->Ele[i].QSS = (urand() - 0.3) * 0.5; }
            if (i % 11 == 5) { MD->Ele[i].iSS = -1; MD->Ele[i].QSS = (urand() - 0.7) * 0.5; }
        }
    }
    if (has(mutate, "ebc")) {
        fake_tsd(MD->tsd_eyBC, 4);
        fake_tsd(MD->tsd_eqBC, 4);
        for (int c = 1; c <= 4; c++) {
            MD->tsd_eyBC.ts[0][c] = 2.0 + 3.0 * urand();
            MD->tsd_eqBC.ts[0][c] = (urand() - 0.5) * 2.0;
        }
        for (int i = 0; i < Ne; i++) {
            if (MD->Ele[i].iLake > 0) continue;
            if (i % 13 == 4) MD->Ele[i].iBC = 1 + (i / 13) % 4;
            if (i % 17 == 6) MD->Ele[i].iBC = -(1 + (i / 17) % 4);
        }
    }
    if (has(mutate, "rbc")) {
        fake_tsd(MD->tsd_ryBC, 3);
        fake_tsd(MD->tsd_rqBC, 3);
        for (int c = 1; c <= 3; c++) {
            MD->tsd_ryBC.ts[0][c] = 0.2 + 1.0 * urand();
            MD->tsd_rqBC.ts[0][c] = 50.0 * urand();
        }
        for (int i = 0; i < Nr; i++) {
            if (i % 9 == 2) MD->Riv[i].BC = 1 + (i / 9) % 3;
            if (i % 5 == 1) MD->Riv[i].BC = -(1 + (i / 5) % 3);
        }
    }
    if (has(mutate, "down4") && !lakeon) {
        for (int i = 0; i < Nr; i++)
            if (MD->Riv[i].down < 0) { MD->Riv[i].down = -4; break; }
    }

    /* ---- state ---- */
    if (state.rfind("rand:", 0) == 0) {
        for (int i = 0; i < Ne; i++) {
            const double aqd = MD->Ele[i].AquiferDepth;
            double r = urand();
            double ysf = (r < 0.45) ? 0. : (r < 0.5 ? -1e-4 * urand() : 0.02 * urand());
            if (urand() < 0.03) ysf = 0.1 + 0.6 * urand(); /* deep ponding: MAXYSURF clamp */
            double ygw = (0.2 + 0.75 * urand()) * aqd;
            r = urand();
            if (r < 0.05) ygw = (1.0 + 0.05 * urand()) * aqd; /* water table above ground */
            else if (r < 0.08) ygw = 0.02 * urand();          /* nearly dry aquifer */
            else if (r < 0.10) ygw = -0.01 * urand();
            double def = aqd - ygw;
            double yus = (def > 0 ? def : 0.01) * (0.05 + 0.55 * urand());
            r = urand();
            if (r < 0.05) yus = (def > 0 ? def : 0.01) * (1.0 + 0.1 * urand()); /* over-full unsat */
            else if (r < 0.10) yus = 0.004 * urand();                           /* <= EPSILON */
            else if (r < 0.12) yus = -1e-3 * urand();
            Y[i] = ysf;
            Y[i + Ne] = yus;
            Y[i + 2 * Ne] = ygw;
        }
        for (int i = 0; i < Nr; i++) {
            double r = urand();
            double d = MD->Riv[i].depth;
            Y[3 * Ne + i] = (r < 0.1) ? 0. : (r < 0.15 ? -1e-3 * urand() : (r < 0.25 ? d * (1.0 + 0.3 * urand()) : 0.5 * d * urand()));
        }
        for (int i = 0; i < Nl; i++) Y[3 * Ne + Nr + i] *= (0.5 + urand());
        /* interception evaporation: ET() leaves 0 when the canopy store is empty (it is, at the
         * initial condition); give f_etFlux both of its branches (qEleE_IC >= / < qPotTran,
         * src/ModelData/MD_ET.cpp:368-379) */
        for (int i = 0; i < Ne; i++)
            MD->qEleE_IC[i] = (urand() < 0.3) ? 0. : 1.5 * urand() * MD->qPotTran[i];
    }
    /* the reference never refreshes uYgw of flux-BC cells (iBC<0,
     * src/ModelData/MD_update.cpp:114-124) - a stale read; the replacement uses
     * Y[iGW], so give the reference the same value to be comparable. */
    for (int i = 0; i < Ne; i++)
        if (MD->Ele[i].iBC < 0) uYgw[i] = Y[i + 2 * Ne];

    g_out = fopen(outfn.c_str(), "wb");
    if (!g_out) { perror("open out"); return 1; }

    /* ---- call #1 (establishes the carried state for this y); its own inputs and
     * result are kept too: the state fresh from updateforcing()+ET() ---- */
    std::vector<double> satn_first(Ne), eic_first(MD->qEleE_IC, MD->qEleE_IC + Ne);
    for (int i = 0; i < Ne; i++) satn_first[i] = MD->Ele[i].u_satn;
    f(t, udata, du, MD);
    std::vector<double> ydot_first(DY, DY + NY);
    {
        double s1 = 0, sa1 = 0;
        for (int i = 0; i < NY; i++) { s1 += DY[i]; sa1 += fabs(DY[i]); }
        printf("\n[shud_ref] first call: sum(ydot)=%.17g sum|ydot|=%.17g\n", s1, sa1);
    }

    /* ---- static data, SoA ---- */
    put1i("Ne", Ne); put1i("Nr", Nr); put1i("Ns", Ns); put1i("Nl", Nl);
    put1i("close_boundary", MD->CS.CloseBoundary);
    put1i("lakeon", lakeon);
    put1d("t", t);
    ELE_D(x) ELE_D(y)
    ELE_D(area) ELE_D(z_surf) ELE_D(z_bottom) ELE_D(depression)
    ELE_D(AquiferDepth) ELE_D(Sy) ELE_D(infD) ELE_D(infKsatV) ELE_D(macKsatV) ELE_D(hAreaF)
    ELE_D(ThetaS) ELE_D(ThetaR) ELE_D(ThetaFC) ELE_D(Alpha) ELE_D(Beta)
    ELE_D(KsatH) ELE_D(KsatV) ELE_D(macKsatH) ELE_D(macD) ELE_D(geo_vAreaF)
    ELE_D(VegFrac) ELE_D(ImpAF) ELE_D(WetlandLevel) ELE_D(RootReachLevel) ELE_D(Rough) ELE_D(QSS)
    ELE_D3(edge) ELE_D3(Dist2Nabor) ELE_D3(Dist2Edge) ELE_D3(avgRough)
    ELE_I3(nabr) ELE_I3(lakenabr)
    ELE_I(iLake) ELE_I(iBC) ELE_I(iSS)
    {
        /* f_etFlux reads ThetaS/ThetaR from the soil table, not the element copy
         * (src/ModelData/MD_ET.cpp:347); check they are the same numbers. */
        for (int i = 0; i < Ne; i++) {
            const Soil_Layer &s = MD->Soil[MD->Ele[i].iSoil - 1];
            if (s.ThetaS != MD->Ele[i].ThetaS || s.ThetaR != MD->Ele[i].ThetaR) {
                fprintf(stderr, "soil table / element copy mismatch at cell %d\n", i + 1);
                return 3;
            }
        }
    }
    RIV_D(Length) RIV_D(BedSlope) RIV_D(depth) RIV_D(BottomWidth) RIV_D(bankslope)
    RIV_D(avgRough) RIV_D(Dist2DownStream) RIV_D(KsatH) RIV_D(BedThick) RIV_D(zbank)
    RIV_I(down) RIV_I(BC) RIV_I(toLake)
    {
        std::vector<int> ie(Ns), ir(Ns);
        std::vector<double> len(Ns), cwr(Ns);
        for (int i = 0; i < Ns; i++) {
            ie[i] = MD->RivSeg[i].iEle; ir[i] = MD->RivSeg[i].iRiv;
            len[i] = MD->RivSeg[i].length; cwr[i] = MD->RivSeg[i].Cwr;
        }
        puti("seg_iEle", ie); puti("seg_iRiv", ir); putd("seg_length", len); putd("seg_Cwr", cwr);
    }
    {
        std::vector<double> zmin(Nl), yi, ai;
        std::vector<int> nele(Nl), ptr(Nl + 1, 0);
        for (int l = 0; l < Nl; l++) {
            zmin[l] = MD->lake[l].zmin; nele[l] = MD->lake[l].NumEleLake;
            for (int k = 0; k < MD->lake[l].bathymetry.nvalue; k++) {
                yi.push_back(MD->lake[l].bathymetry.yi[k]);
                ai.push_back(MD->lake[l].bathymetry.ai[k]);
            }
            ptr[l + 1] = (int)yi.size();
        }
        putd("lake_zmin", zmin); puti("lake_NumEleLake", nele);
        puti("lake_bathy_ptr", ptr); putd("lake_bathy_yi", yi); putd("lake_bathy_ai", ai);
    }
    /* ---- values fixed between forcing steps ---- */
    putd("qEleNetPrep", MD->qEleNetPrep, Ne); putd("qPotEvap", MD->qPotEvap, Ne);
    putd("qPotTran", MD->qPotTran, Ne); putd("t_lai", MD->t_lai, Ne);
    putd("fu_Surf", MD->fu_Surf, Ne); putd("fu_Sub", MD->fu_Sub, Ne);
    putd("qElePrep", MD->qElePrep, Ne); putd("qEleETP", MD->qEleETP, Ne);
    ELE_D(yBC) ELE_D(QBC) RIV_D(yBC) RIV_D(qBC)
    /* ---- carried state before call #2 ---- */
    ELE_D(u_satn)
    putd("qEleE_IC_in", MD->qEleE_IC, Ne);
    putd("y", Y, NY);
    putd("first_u_satn", satn_first); putd("first_qEleE_IC_in", eic_first); putd("first_ydot", ydot_first);

    /* ---- call #2: the outputs that pin everything ---- */
    f(t, udata, du, MD);
    putd("ydot", DY, NY);
    putd("qEleE_IC_out", MD->qEleE_IC, Ne);
    {
        std::vector<double> v(Ne);
        for (int i = 0; i < Ne; i++) v[i] = MD->Ele[i].u_satn;
        putd("u_satn_out", v);
        for (int i = 0; i < Ne; i++) v[i] = MD->Ele[i].u_effKH;
        putd("u_effKH", v);
        for (int i = 0; i < Ne; i++) v[i] = MD->Ele[i].u_satKr;
        putd("u_satKr", v);
    }
    putd("qEleInfil", MD->qEleInfil, Ne); putd("qEleExfil", MD->qEleExfil, Ne);
    putd("qEleRecharge", MD->qEleRecharge, Ne);
    putd("qEs", MD->qEs, Ne); putd("qEu", MD->qEu, Ne); putd("qEg", MD->qEg, Ne);
    putd("qTu", MD->qTu, Ne); putd("qTg", MD->qTg, Ne);
    putd("qEleTrans", MD->qEleTrans, Ne); putd("qEleEvapo", MD->qEleEvapo, Ne);
    putd("qEleETA", MD->qEleETA, Ne); putd("iBeta", MD->iBeta, Ne);
    {
        std::vector<double> a(3 * (size_t)Ne), b(3 * (size_t)Ne);
        for (int j = 0; j < 3; j++)
            for (int i = 0; i < Ne; i++) {
                a[(size_t)j * Ne + i] = MD->QeleSurf[i][j];
                b[(size_t)j * Ne + i] = MD->QeleSub[i][j];
            }
        putd("QeleSurf", a); putd("QeleSub", b);
    }
    putd("QeleSurfTot", MD->QeleSurfTot, Ne); putd("QeleSubTot", MD->QeleSubTot, Ne);
    putd("Qe2r_Surf", MD->Qe2r_Surf, Ne); putd("Qe2r_Sub", MD->Qe2r_Sub, Ne);
    putd("QsegSurf", MD->QsegSurf, Ns); putd("QsegSub", MD->QsegSub, Ns);
    putd("QrivSurf", MD->QrivSurf, Nr); putd("QrivSub", MD->QrivSub, Nr);
    putd("QrivUp", MD->QrivUp, Nr); putd("QrivDown", MD->QrivDown, Nr);
    if (Nl > 0) {
        putd("y2LakeArea", MD->y2LakeArea, Nl); putd("QLakeSurf", MD->QLakeSurf, Nl);
        putd("QLakeSub", MD->QLakeSub, Nl); putd("QLakeRivIn", MD->QLakeRivIn, Nl);
        putd("QLakeRivOut", MD->QLakeRivOut, Nl); putd("qLakeEvap", MD->qLakeEvap, Nl);
        putd("qLakePrcp", MD->qLakePrcp, Nl);
    }
    if (fseq > 0) {
        const char *names[] = {"qEleNetPrep", "qPotEvap", "qPotTran", "t_lai", "qEleE_IC", "qElePrep", "fu_Surf", "fu_Sub"};
        std::vector<std::vector<double>> seq(8);
        std::vector<double> tt;
        double tf = MD->CS.StartTime;
        /* a fresh model: the buckets must start from the initial condition */
        Model_Data *M2 = new Model_Data(fin, fout);
        M2->loadinput(); M2->initialize(); M2->CheckInputData(); M2->LoadIC();
        for (int k = 0; k < fseq; k++, tf += 60.) {
            M2->updateAllTimeSeries(tf);
            M2->updateforcing(tf);
            M2->ET(tf, tf + 60.);
            const double *src[] = {M2->qEleNetPrep, M2->qPotEvap, M2->qPotTran, M2->t_lai, M2->qEleE_IC, M2->qElePrep,
                                   M2->fu_Surf, M2->fu_Sub};
            for (int a = 0; a < 8; a++) seq[a].insert(seq[a].end(), src[a], src[a] + Ne);
            tt.push_back(tf);
        }
        for (int a = 0; a < 8; a++) putd((std::string("fseq_") + names[a]).c_str(), seq[a]);
        putd("fseq_t", tt);
    }
    if (!print_init.empty()) {
        for (int i = 0; i < Ne; i++) {
            MD->yEleIS[i] = (urand() < 0.3) ? 0. : 2e-3 * urand();
            MD->yEleSnow[i] = (urand() < 0.5) ? 0. : 0.4 * urand();
        }
        putd("ic_yEleIS", MD->yEleIS, Ne); putd("ic_yEleSnow", MD->yEleSnow, Ne);
        put1d("ic_t", t + 1440.);
        MD->summary(udata);
        MD->CS.UpdateICStep = 1;
        MD->PrintInit(print_init.c_str(), t + 1440.);
    }
    if (lseq > 0) {
        Model_Data *M3 = new Model_Data(fin, fout);
        M3->loadinput(); M3->initialize(); M3->CheckInputData(); M3->LoadIC();
        /* frozen-soil factors (CS.cryosphere, MD_ET.cpp:301-311): no shipped basin switches them on.  The
         * accumulators' ACC member has no initialiser (AccTemperature.hpp:24): start it at 0 explicitly. */
        if (has(mutate, "cryo")) {
            M3->CS.cryosphere = 1;
            M3->gc.cTemp = -6.0; /* calibration offset: ccw's winter is too mild to freeze anything otherwise */
            for (int i = 0; i < Ne; i++) { M3->AccT_surf[i].ACC = 0.; M3->AccT_sub[i].ACC = 0.; }
        }
        const int nf = M3->NumForc;
        int nlc = 0, nmf = 0;
        {
            std::vector<int> iForc(Ne), iLC(Ne), iMF(Ne);
            std::vector<double> alb(Ne), fp(Ne), wh(Ne), nx(Ne), ny(Ne), nz(Ne);
            for (int i = 0; i < Ne; i++) {
                iForc[i] = M3->Ele[i].iForc; iLC[i] = M3->Ele[i].iLC; iMF[i] = M3->Ele[i].iMF;
                alb[i] = M3->Ele[i].Albedo; fp[i] = M3->Ele[i].FixPressure; wh[i] = M3->Ele[i].windH;
                nx[i] = M3->Ele[i].nx; ny[i] = M3->Ele[i].ny; nz[i] = M3->Ele[i].nz;
                nlc = std::max(nlc, iLC[i]); nmf = std::max(nmf, iMF[i]);
            }
            puti("land_iForc", iForc); puti("land_iLC", iLC); puti("land_iMF", iMF);
            putd("land_Albedo", alb); putd("land_FixPressure", fp); putd("land_windH", wh);
            putd("land_nx", nx); putd("land_ny", ny); putd("land_nz", nz);
            std::vector<double> fz(nf);
            for (int k = 0; k < nf; k++) fz[k] = M3->forcing ? M3->forcing->z(k) : M3->tsd_weather[k].xyz[2];
            putd("land_forc_z", fz);
            const double gcv[] = {M3->gc.cPrep, M3->gc.cTemp, M3->gc.cLAItsd, M3->gc.cMF, M3->gc.cETP, M3->gc.cISmax};
            putd("land_gc", gcv, 6);
            const double csv[] = {(double)M3->CS.radiation_input_mode, (double)M3->CS.terrain_radiation, (double)M3->CS.cryosphere,
                                  M3->CS.rad_factor_cap, M3->CS.rad_cosz_min, (double)SWNET};
            putd("land_cs", csv, 6);
            putd("land_yEleSnow0", M3->yEleSnow, Ne); putd("land_yEleIS0", M3->yEleIS, Ne);
            const double frz[] = {(double)M3->AccT_surf[0].MaxLen, M3->AccT_surf_max, M3->AccT_surf_min,
                                 (double)M3->AccT_sub[0].MaxLen, M3->AccT_sub_max, M3->AccT_sub_min};
            putd("land_frozen", frz, 6);
        }
        const char *onames[] = {"qElePrep", "qPotEvap", "qPotTran", "qEleETP", "t_lai", "t_temp", "t_mf", "qEleNetPrep",
                                "qEleE_IC", "yEleSnow", "yEleIS", "fu_Surf", "fu_Sub", "rn_factor"};
        std::vector<std::vector<double>> out(14);
        std::vector<double> tt, frows, lai, mf, sx, sy, sz, wdt, den;
        std::vector<int> sn, kept;
        double tf = std::isnan(land_t0) ? M3->CS.StartTime : land_t0;
        for (int k = 0; k < lseq; k++, tf += land_dt) {
            M3->updateAllTimeSeries(tf);
            M3->updateforcing(tf);
            M3->ET(tf, tf + land_dt);
            for (int st = 0; st < nf; st++)
                for (int col = 1; col <= 5; col++)
                    frows.push_back(M3->forcing ? M3->forcing->get(st, col) : M3->tsd_weather[st].getX(tf, col));
            for (int c = 1; c <= nlc; c++) lai.push_back(M3->tsd_LAI.getX(tf, c));
            for (int c = 1; c <= nmf; c++) mf.push_back(M3->tsd_MF.getX(tf, c));
            sn.push_back(M3->tsr_forcing_n);
            den.push_back(M3->tsr_forcing_den);
            sx.insert(sx.end(), M3->tsr_forcing_sx.begin(), M3->tsr_forcing_sx.end());
            sy.insert(sy.end(), M3->tsr_forcing_sy.begin(), M3->tsr_forcing_sy.end());
            sz.insert(sz.end(), M3->tsr_forcing_sz.begin(), M3->tsr_forcing_sz.end());
            wdt.insert(wdt.end(), M3->tsr_forcing_wdt.begin(), M3->tsr_forcing_wdt.end());
            const double *src[] = {M3->qElePrep, M3->qPotEvap, M3->qPotTran, M3->qEleETP, M3->t_lai, M3->t_temp, M3->t_mf,
                                   M3->qEleNetPrep, M3->qEleE_IC, M3->yEleSnow, M3->yEleIS, M3->fu_Surf, M3->fu_Sub,
                                   M3->ele_rn_factor};
            if (k % lstride == lstride - 1 || k == lseq - 1) {  /* outputs kept every lstride-th step; inputs always */
                for (int a = 0; a < 14; a++) out[a].insert(out[a].end(), src[a], src[a] + Ne);
                kept.push_back(k);
            }
            tt.push_back(tf);
        }
        put1i("land_nforc", nf); put1i("land_nlc", nlc); put1i("land_nmf", nmf);
        put1d("lseq_dt", land_dt);
        putd("lseq_t", tt); putd("lseq_forc", frows); putd("lseq_lai", lai); putd("lseq_mf", mf);
        puti("lseq_tsr_n", sn); putd("lseq_tsr_den", den); puti("lseq_kept", kept);
        putd("lseq_tsr_sx", sx); putd("lseq_tsr_sy", sy); putd("lseq_tsr_sz", sz); putd("lseq_tsr_wdt", wdt);
        for (int a = 0; a < 14; a++) putd((std::string("lseq_") + onames[a]).c_str(), out[a]);
    }
    fclose(g_out);

    double s = 0, sa = 0;
    for (int i = 0; i < NY; i++) { s += DY[i]; sa += fabs(DY[i]); }
    printf("\n[shud_ref] %s Ne=%d Nr=%d Ns=%d Nl=%d NY=%d t=%.1f sum(ydot)=%.17g sum|ydot|=%.17g\n",
           prj.c_str(), Ne, Nr, Ns, Nl, NY, t, s, sa);

    if (reps > 0) {
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; r++) f(t, udata, du, MD);
        auto t1 = std::chrono::steady_clock::now();
        double us = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
        printf("[shud_ref] time_per_f_us=%.3f cells_per_s=%.6g reps=%d\n", us, Ne / (us * 1e-6), reps);
    }
    return 0;
}
