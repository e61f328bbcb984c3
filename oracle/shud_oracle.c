/* TEST INFRASTRUCTURE ONLY - never linked into, imported by or executed from the
 * product path (shud_up_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it, and only as the checker / CPU baseline.
 *
 * Plain-C restatement of the reference's serial RHS f(t,y,ydot) on SoA inputs
 * (the same shud_mesh / shud_forcing structs the C ABI takes, include/shud_b200.h).
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  It is written from the reference's *behaviour*: same branches,
 * same literal constants, same floating-point evaluation order, same summation order.
 *
 * PINNED: tests/test_oracle_golden.py checks it bit-for-bit (ydot and every flux array)
 * against snapshots of the unmodified reference f() compiled from /root/reference by
 * oracle/Makefile + oracle/ref_driver.cpp (tests/golden/ *.npz: ccw, heihe, qhh at the
 * initial condition and at randomised states, plus in-memory mutations that switch on the
 * BC / SS / open-boundary / frozen-soil / critical-depth branches no shipped basin reaches).
 * Build with -ffp-contract=off: the x86-64 reference build emits no FMA.
 *
 * Deliberate deviation (documented): for flux-BC cells (iBC<0) the
 * reference never refreshes uYgw[i] (src/ModelData/MD_update.cpp:114-124, a stale read);
 * here uYgw[i] = Y[iGW].  oracle/ref_driver.cpp feeds the reference the same value.
 */
#include "shud_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* src/Model/Macros.hpp:31-35,46,51,67 - literal values kept, not corrected */
#define EPSILON 0.005
#define ZERO 1.0e-10
#define EPS_SLOPE 0.05e-6
#define PI 3.1415926
#define GRAV 9.8
#define MAXYSURF 0.5
#define NA_VALUE -9999
#define FieldCapacityRatio 0.75
/* src/Equations/functions.hpp:117-123 (NaN semantics differ from fmin/fmax) */
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#define MAX(a, b) ((a) < (b) ? (b) : (a))

#include <stdio.h>
#include <time.h>
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static int g_prof = -1;
#define PROF(tag) do { if (g_prof > 0) { double t_ = now_s(); fprintf(stderr, "[oracle] %-10s %.1f ms\n", tag, (t_ - t_prof) * 1e3); t_prof = t_; } } while (0)
const char *shud_oracle_version(void) { return "shud_oracle r1 (serial f(), SoA)"; }

/* src/Equations/functions.cpp:90-103 */
static int bad_nan(double x) { return isnan(x) || isinf(x); }
/* src/Equations/functions.cpp:148-154 */
static int bad_nonneg(double x) { return x < 0.0 || isnan(x) || isinf(x) || fabs(x - NA_VALUE) < ZERO; }

/* src/Equations/Equations.hpp:36-39 */
static double pow23(double x) {
    double t = cbrt(x);
    return t * t;
}
/* src/Equations/Equations.hpp:54-63 */
static double ManningEquation(double Area, double rough, double R, double S) {
    if (S > 0) {
        return sqrt(S) * Area * pow23(R) / rough;
    } else {
        return -1.0 * sqrt(-S) * Area * pow23(R) / rough;
    }
}
/* src/Equations/Equations.hpp:45-48 */
static double meanHarmonic(double k1, double k2, double d1, double d2) {
    return (k1 * k2) * (d1 + d2) / (d1 * k2 + d2 * k1);
}
/* src/Equations/Equations.hpp:50-52 */
static double meanArithmetic(double k1, double k2, double d1, double d2) {
    return (k1 * d1 + k2 * d2) / (d1 + d2);
}
/* src/Equations/Equations.cpp:8-51 */
static double avgY_sf(double z1, double y1, double z2, double y2, double threshold) {
    double h1 = z1 + y1, h2 = z2 + y2;
    if (h1 > h2) {
        return (y1 > threshold) ? y1 : 0.;
    } else {
        return (y2 > threshold) ? y2 : 0.;
    }
}
/* src/Equations/Equations.cpp:52-70 */
static double avgY_gw(double y1, double y2) {
    y1 = MAX(y1, 0.);
    y2 = MAX(y2, 0.);
    return (y1 + y2) * .5;
}
/* src/Equations/Equations.cpp:116-134; *err = 13 where the reference calls myexit(ERRDATAIN) */
static double effKH(double Ygw, double aqDepth, double MacD, double Kmac, double AF, double Kmx, int *err) {
    double effk = 0;
    if (MacD <= ZERO || Ygw < aqDepth - MacD) {
        effk = Kmx;
    } else {
        if (Ygw > aqDepth) {
            effk = (Kmac * MacD * AF + Kmx * (aqDepth - MacD * AF)) / aqDepth;
        } else {
            effk = (Kmac * (Ygw - (aqDepth - MacD)) * AF +
                    Kmx * (aqDepth - MacD + (Ygw - (aqDepth - MacD)) * (1 - AF))) / Ygw;
        }
    }
    if (effk < 0. || effk > 1e9) *err = SHUD_ERRDATAIN;
    return effk;
}
/* src/Equations/Equations.cpp:136-141 */
static double satKfun(double elemSatn, double n) {
    double temp = -1. + pow(1. - pow(elemSatn, n / (n - 1.)), (n - 1.) / n);
    double ret = sqrt(elemSatn) * temp * temp;
    return ret;
}
/* src/Equations/is_sm_et.cpp:131-142 */
static double SoilMoistureStress(double ThetaS, double ThetaR, double SatRatio) {
    double fc, beta_s;
    fc = ThetaS * FieldCapacityRatio;
    beta_s = (SatRatio * (ThetaS - ThetaR) - ThetaR) / (fc - ThetaR);
    beta_s = MIN(MAX(0., beta_s), 1.);
    beta_s = 0.5 * (1 - cos(PI * beta_s));
    return beta_s;
}
/* src/ModelData/MD_RiverFlux.cpp:65-98 */
static double WeirFlow_jtoi(double zi, double yi, double zj, double yj, double zbank, double cwr, double width,
                            double threshold) {
    double hi, hj, Q = 0.;
    double dh, y;
    hi = yi + zi;
    hj = yj + zj;
    dh = hj - hi;
    if (dh > 0.) {
        y = hi - zbank;
        if ((y > 0.) & (yj > threshold)) {
            if (hi > zbank) {
                y = dh;
            }
            Q = cwr * sqrt(2. * GRAV * y) * width * y * 60.;
        } else {
            Q = 0.;
        }
    } else {
        y = hi - zbank;
        if (y > 0. && yi > threshold) {
            if (hj > zbank) {
                y = -dh;
            }
            Q = -1. * cwr * sqrt(2. * GRAV * y) * width * y * 60.;
        } else {
            Q = 0.;
        }
    }
    return Q;
}
/* src/Equations/Flux_RiverElement.cpp:11-55 */
static double flux_R2E_GW(double yr, double zr, double ye, double ze, double Kele, double Kriv, double L,
                          double D_riv) {
    double dh, A, g, K, he, hr;
    double Q = 0.0;
    if (Kele < ZERO || Kriv < ZERO) {
        return 0.;
    } else {
        K = meanArithmetic(Kele, Kriv, 1., 1.);
    }
    he = ye + ze;
    hr = yr + zr;
    dh = hr - he;
    if (dh > ZERO) {
        if (he > zr) {
            A = (yr + (he - zr)) * .5 * L;
        } else {
            A = yr * L;
        }
        if (yr < EPSILON) {
            Q = 0.;
        } else {
            g = dh / D_riv;
            Q = A * K * g;
        }
    } else if (dh < -ZERO) {
        if (ye > ZERO) {
            A = (yr + (he - zr)) * .5 * L;
            g = dh / D_riv;
            Q = A * K * g;
        } else {
            Q = 0.;
        }
    } else {
        Q = 0.;
    }
    return Q;
}
/* src/Equations/functions.hpp:125-139 */
static double Quadratic(double s, double w, double dA) {
    double ret = 0., cc;
    s = fabs(s);
    cc = w * w + 4 * s * dA;
    if (cc < ZERO) {
        ret = -1. * w / (2. * s);
    } else {
        ret = (-w + sqrt(cc)) / (2 * s);
    }
    return ret;
}
/* src/Equations/functions.hpp:141-153 */
static double fun_dAtodY(double dA, double w_top, double s) {
    double dy = 0.;
    if (dA == 0.) return 0.;
    if (fabs(s) < EPS_SLOPE) {
        dy = dA / w_top;
    } else {
        dy = Quadratic(s, w_top, dA);
    }
    return dy;
}
/* src/classes/Lake.cpp:59-78 (slope uses yi[i]-y, as written there) */
static double lake_toparea(const double *yi, const double *ai, int nvalue, double y) {
    double ta = ai[0];
    double dy, da;
    if (y <= yi[0]) {
        ta = ai[0];
    } else {
        for (int i = 1; i < nvalue; i++) {
            if (y < yi[i]) {
                da = (ai[i] - ta);
                dy = yi[i] - y;
                ta = da / dy * (y - yi[i - 1]) + ta;
                break;
            } else {
                ta = ai[i];
            }
        }
    }
    return ta;
}
static double fixMaxValue(double x, double defVal) { return (x < defVal) ? defVal : x; } /* functions.hpp:183-189 */

/* per-cell cached members of _Element that survive inside one call (src/classes/Element.hpp:101-116) */
typedef struct {
    double *effKH, *deficit, *Kmax, *satn, *theta, *satKr;
} cellcache;

/* _Element::updateElement, src/classes/Element.cpp:347-384 (u_phius / u_effkInfi there are dead stores) */
static void updateElement(const shud_mesh *m, int i, double Yunsat, double Ygw, cellcache *c, int *err) {
    double u_deficit, u_satn, u_theta, u_satKr;
    c->effKH[i] = effKH(Ygw, m->AquiferDepth[i], m->macD[i], m->macKsatH[i], m->geo_vAreaF[i], m->KsatH[i], err);
    u_deficit = m->AquiferDepth[i] - Ygw;
    c->Kmax[i] = m->infKsatV[i] * (1. - m->hAreaF[i]) + m->macKsatV[i] * m->hAreaF[i];
    if (u_deficit <= 0.) {
        u_deficit = 0.;
        u_satn = 1.;
        u_theta = m->ThetaS[i];
    } else {
        u_theta = Yunsat / u_deficit * m->ThetaS[i];
        u_satn = (u_theta - m->ThetaR[i]) / (m->ThetaS[i] - m->ThetaR[i]);
    }
    if (u_satn > 0.99) {
        u_satn = 1.0;
        u_satKr = 1.0;
        u_theta = m->ThetaS[i];
    } else if (u_satn <= ZERO) {
        u_satn = 0.;
        u_satKr = 0.;
        u_theta = m->ThetaR[i];
    } else {
        u_satKr = satKfun(u_satn, m->Beta[i]);
    }
    c->deficit[i] = u_deficit;
    c->satn[i] = u_satn;
    c->theta[i] = u_theta;
    c->satKr[i] = u_satKr;
}

void shud_oracle_prime(const shud_mesh *m, const double *y, double *u_satn) {
    const int Ne = m->Ne;
    cellcache c;
    double *buf = (double *)malloc(sizeof(double) * 5 * (size_t)Ne);
    c.effKH = buf; c.deficit = buf + Ne; c.Kmax = buf + 2 * (size_t)Ne; c.theta = buf + 3 * (size_t)Ne;
    c.satKr = buf + 4 * (size_t)Ne; c.satn = u_satn;
    int err = 0;
    /* Model_Data::updateforcing calls updateElement(uYsf,uYus,uYgw) for EVERY cell (lake cells too),
     * src/ModelData/MD_ET.cpp:14-19; uYgw of a head-BC cell would be yBC - priming ignores BCs. */
    for (int i = 0; i < Ne; i++) updateElement(m, i, y[i + Ne], y[i + 2 * (size_t)Ne], &c, &err);
    free(buf);
}

#define STORE(arr, idx, val) do { if (diag && diag->arr) diag->arr[idx] = (val); } while (0)

int shud_oracle_rhs(const shud_mesh *m, const shud_forcing *F, double *u_satn_io, double *qEleE_IC,
                    const double *Y, double *DY, const shud_diag *diag, int nthreads) {
    const int Ne = m->Ne, Nr = m->Nr, Ns = m->Ns, Nl = m->Nl;
    const size_t NE = (size_t)Ne;
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#ifndef _OPENMP
    nthreads = 1;
#endif
    if (g_prof < 0) g_prof = getenv("SHUD_ORACLE_PROF") ? 1 : 0;
    double t_prof = now_s();
    /* ---- scratch: the reference's global/Model_Data work arrays.  Like the reference's, they
     * persist between calls (one cached workspace per process; not thread-safe - neither is f()). ---- */
    size_t nd = 26 * NE + 12 * (size_t)Nr + 2 * (size_t)Ns + 9 * (size_t)Nl + 16;
    static double *W = NULL;
    static size_t Wn = 0;
    if (Wn < nd) {
        free(W);
        W = (double *)malloc(nd * sizeof(double));
        Wn = W ? nd : 0;
        if (!W) return SHUD_ERR_ARG;
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) schedule(static)
        for (long k = 0; k < (long)nd; k++) W[k] = 0.; /* first touch spread over the threads */
    }
    double *p = W;
#define TAKE(n) (p += (n), p - (n))
    double *uYsf = TAKE(NE), *uYus = TAKE(NE), *uYgw = TAKE(NE), *QBC = TAKE(NE);
    double *qInfil = TAKE(NE), *qExfil = TAKE(NE), *qRech = TAKE(NE);
    double *qEs = TAKE(NE), *qEu = TAKE(NE), *qEg = TAKE(NE), *qTu = TAKE(NE), *qTg = TAKE(NE);
    double *qEvapo = TAKE(NE);
    cellcache c;
    c.effKH = TAKE(NE); c.deficit = TAKE(NE); c.Kmax = TAKE(NE); c.theta = TAKE(NE); c.satKr = TAKE(NE);
    c.satn = u_satn_io;
    double *QeleSurf = TAKE(3 * NE), *QeleSub = TAKE(3 * NE);
    double *Qe2rS = TAKE(NE), *Qe2rG = TAKE(NE);
    double *uYriv = TAKE(Nr), *qBC = TAKE(Nr), *topWidth = TAKE(Nr), *CSarea = TAKE(Nr), *CSperem = TAKE(Nr);
    double *QrivSurf = TAKE(Nr), *QrivSub = TAKE(Nr), *QrivUp = TAKE(Nr), *QrivDown = TAKE(Nr);
    double *QsegSurf = TAKE(Ns), *QsegSub = TAKE(Ns);
    double *yLakeStg = TAKE(Nl), *y2LakeArea = TAKE(Nl), *QLakeSurf = TAKE(Nl), *QLakeSub = TAKE(Nl);
    double *QLakeRivIn = TAKE(Nl), *QLakeRivOut = TAKE(Nl), *qLakeEvap = TAKE(Nl), *qLakePrcp = TAKE(Nl);
    /* the accumulators f_update resets every call (MD_update.cpp:165-185) */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) schedule(static)
    for (int i = 0; i < Ne; i++) { Qe2rS[i] = 0.; Qe2rG[i] = 0.; }
    for (int i = 0; i < Nr; i++) { QrivSurf[i] = 0.; QrivSub[i] = 0.; QrivUp[i] = 0.; }
    for (int l = 0; l < Nl; l++) {
        QLakeSurf[l] = 0.; QLakeSub[l] = 0.; qLakeEvap[l] = 0.; qLakePrcp[l] = 0.; QLakeRivIn[l] = 0.; QLakeRivOut[l] = 0.;
    }

    PROF("alloc");
    /* ================= f_update, src/ModelData/MD_update.cpp:102-189 ================= */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) schedule(static)
    for (int i = 0; i < Ne; i++) {
        uYsf[i] = Y[i];
        uYus[i] = Y[i + NE];
        if (m->iBC[i] == 0) {
            uYgw[i] = Y[i + 2 * NE];
            QBC[i] = 0.;
        } else if (m->iBC[i] > 0) {
            uYgw[i] = F->ele_yBC ? F->ele_yBC[i] : 0.;
            QBC[i] = 0.;
        } else {
            uYgw[i] = Y[i + 2 * NE]; /* deviation: see header */
            QBC[i] = F->ele_QBC ? F->ele_QBC[i] : 0.;
        }
    }
    for (int i = 0; i < Nr; i++) {
        uYriv[i] = Y[3 * NE + i];
        /* _River::updateRiver, src/classes/River.cpp:49-62 + River.hpp:115-128 (only the members f() reads) */
        const double y = uYriv[i], w0 = m->riv_BottomWidth[i], s = m->riv_bankslope[i];
        topWidth[i] = fixMaxValue(y * s * 2.0 + w0, 0.);
        CSarea[i] = fixMaxValue(y * (w0 + y * s), 0.);
        CSperem[i] = fixMaxValue(2.0 * sqrt(y * y + (y * s) * (y * s)) + w0, 0.);
        qBC[i] = 0.0;
        if (m->riv_BC[i] < 0) {
            qBC[i] = F->riv_qBC ? F->riv_qBC[i] : 0.;
        } else if (m->riv_BC[i] > 0) {
            uYriv[i] = F->riv_yBC ? F->riv_yBC[i] : 0.;
        }
    }
    for (int l = 0; l < Nl; l++) {
        yLakeStg[l] = Y[3 * NE + Nr + l];
        const int b0 = m->lake_bathy_ptr[l], b1 = m->lake_bathy_ptr[l + 1];
        /* _Lake::update, src/classes/Lake.cpp:104-107 */
        y2LakeArea[l] = lake_toparea(m->lake_bathy_yi + b0, m->lake_bathy_ai + b0, b1 - b0, yLakeStg[l] + m->lake_zmin[l]);
    }

    PROF("f_update");
    /* ================= f_loop, src/ModelData/MD_f.cpp:9-50 ================= */
    /* ---- LOOP A (MD_f.cpp:11-26) ---- */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) reduction(max : err) schedule(static)
    for (int i = 0; i < Ne; i++) {
        int e = 0;
        if (m->lakeon && m->iLake[i] > 0) {
            /* _Element::updateLakeElement (Element.cpp:336-346) + fun_Ele_lakeVertical (MD_ElementFlux.cpp:2-17) */
            c.effKH[i] = m->KsatH[i];
            c.deficit[i] = 0.; c.Kmax[i] = m->infKsatV[i]; c.satn[i] = 1.; c.theta[i] = m->ThetaS[i]; c.satKr[i] = 1.0;
            qInfil[i] = 0.; qRech[i] = 0.; qExfil[i] = 0.;
            qEs[i] = qEu[i] = qEg[i] = qTu[i] = qTg[i] = 0.;
            qEleE_IC[i] = 0.;
            qEvapo[i] = F->qPotEvap[i];
            STORE(qEleTrans, i, 0.); STORE(qEleEvapo, i, qEvapo[i]);
            STORE(qEleETA, i, qEleE_IC[i] + qEvapo[i] + 0.);
            continue;
        }
        /* ---- f_etFlux, src/ModelData/MD_ET.cpp:343-404 ---- */
        {
            double Es = 0., Eu = 0., Tu = 0., Eg = 0., Tg = 0.;
            const double va = m->VegFrac[i], vb = 1. - m->VegFrac[i];
            const double pj = 1. - m->ImpAF[i];
            const double iBeta = SoilMoistureStress(m->ThetaS[i], m->ThetaR[i], c.satn[i]);
            const double pe = F->qPotEvap[i], pt = F->qPotTran[i];
            Es = MIN(MAX(0., uYsf[i]), pe) * vb;
            if (Es < pe) {
                if (uYgw[i] > m->WetlandLevel[i]) {
                    Eg = MIN(MAX(0., uYgw[i]), pe - Es) * pj * vb;
                    Eu = 0.;
                } else {
                    Eg = 0.;
                    Eu = MIN(MAX(0., uYus[i]), iBeta * (pe - Es)) * pj * vb;
                }
            } else {
                Eg = 0.;
                Eu = 0.;
            }
            if (F->t_lai[i] > ZERO) {
                if (qEleE_IC[i] >= pt) {
                    Tg = Tu = 0.;
                    qEleE_IC[i] = pt * pj * va;
                } else {
                    if (uYgw[i] > m->RootReachLevel[i]) {
                        Tg = MIN(MAX(0., uYgw[i]), (pt - qEleE_IC[i])) * pj * va;
                        Tu = 0.;
                    } else {
                        Tg = 0.;
                        Tu = MIN(MAX(0., uYus[i]), iBeta * (pt - qEleE_IC[i])) * pj * va;
                    }
                }
            } else {
                Tg = Tu = qEleE_IC[i] = 0.;
            }
            qEs[i] = Es; qEu[i] = Eu; qEg[i] = Eg; qTu[i] = Tu; qTg[i] = Tg;
            const double trans = Tg + Tu, evapo = Eu + Eg + Es, eta = qEleE_IC[i] + evapo + trans;
            qEvapo[i] = evapo;
            STORE(qEleTrans, i, trans); STORE(qEleEvapo, i, evapo); STORE(qEleETA, i, eta); STORE(iBeta, i, iBeta);
            if (bad_nonneg(Es) || bad_nonneg(Eu) || bad_nonneg(Eg) || bad_nonneg(Tu) || bad_nonneg(Tg) ||
                bad_nan(eta) || bad_nan(evapo) || bad_nan(trans))
                e = SHUD_ERRNAN;
        }
        /* ---- updateElement ---- */
        {
            int e2 = 0;
            updateElement(m, i, uYus[i], uYgw[i], &c, &e2);
            if (e2 && !e) e = e2;
        }
        /* ---- fun_Ele_Infiltraion (MD_ElementFlux.cpp:30-34) -> Flux_Infiltration (Element.cpp:271-303) ---- */
        {
            const double Ysurf = uYsf[i], Yunsat = uYus[i], Ygw = uYgw[i], netprcp = F->qEleNetPrep[i];
            const double AqD = m->AquiferDepth[i], infD = m->infD[i], infKsatV = m->infKsatV[i], hAreaF = m->hAreaF[i],
                         macKsatV = m->macKsatV[i];
            double av = Ysurf + netprcp, grad = 0, u_qex, u_qi, effkInfi;
            if (Ygw + Yunsat > AqD || c.deficit[i] < Yunsat) {
                u_qex = fabs(Ygw + Yunsat - AqD) / AqD * c.Kmax[i];
                u_qi = 0.;
            } else {
                u_qex = 0.;
                if (av > 0. && c.deficit[i] > infD) {
                    grad = 1. + av / infD;
                    if (av > c.Kmax[i]) {
                        effkInfi = infKsatV * (1 - hAreaF) + hAreaF * macKsatV * c.satn[i];
                    } else if (av > infKsatV) {
                        effkInfi = c.satKr[i] * infKsatV * (1 - hAreaF) + hAreaF * macKsatV * c.satn[i];
                    } else {
                        effkInfi = c.satKr[i] * infKsatV * (1 - hAreaF);
                    }
                    u_qi = grad * effkInfi;
                    u_qi = MIN(av, MAX(0., u_qi));
                } else {
                    u_qi = 0;
                }
            }
            qInfil[i] = u_qi * F->fu_Surf[i];
            qExfil[i] = u_qex * F->fu_Surf[i];
        }
        /* ---- fun_Ele_Recharge (MD_ElementFlux.cpp:24-28) -> Flux_Recharge (Element.cpp:304-335) ---- */
        {
            const double Yunsat = uYus[i], Ygw = uYgw[i];
            double ke = 0., grad, ku, u_qr;
            if (Ygw > m->AquiferDepth[i] - m->infD[i] && Yunsat < c.deficit[i]) {
                u_qr = 0.;
            } else {
                if (c.theta[i] > m->ThetaR[i]) {
                    if (Yunsat <= EPSILON) {
                        grad = 0.;
                    } else {
                        grad = (c.theta[i] - m->ThetaR[i]) / (m->ThetaFC[i] - m->ThetaR[i]);
                        grad = MAX(grad, 0.);
                    }
                } else {
                    grad = 0.;
                }
                if (m->infKsatV[i] <= 0. || m->KsatV[i] <= 0.) {
                    u_qr = 0.;
                } else {
                    ku = m->infKsatV[i] * c.satKr[i];
                    ke = meanHarmonic(ku, m->KsatV[i], c.deficit[i], Ygw);
                    u_qr = grad * ke;
                }
            }
            qRech[i] = u_qr * F->fu_Sub[i];
        }
        if (e > err) err = e;
    }
    PROF("loopA");
    /* lake-cell accumulations of LOOP A, ascending cell order (MD_f.cpp:16-17) */
    if (m->lakeon)
        for (int i = 0; i < Ne; i++)
            if (m->iLake[i] > 0) {
                const int l = m->iLake[i] - 1;
                qLakeEvap[l] += qEvapo[i] / m->lake_NumEleLake[l];
                qLakePrcp[l] += F->qElePrep[i] / m->lake_NumEleLake[l];
            }

    /* ---- LOOP B (MD_f.cpp:27-36): fun_Ele_surface / fun_Ele_sub, MD_ElementFlux.cpp:35-156 ---- */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) schedule(static)
    for (int i = 0; i < Ne; i++) {
        if (m->lakeon && m->iLake[i] > 0) {
            for (int j = 0; j < 3; j++) { QeleSurf[j * NE + i] = 0.; QeleSub[j * NE + i] = 0.; }
            continue;
        }
        double isf = uYsf[i];
        isf = isf < 0. ? 0. : isf;
        for (int j = 0; j < 3; j++) {
            const int inabr = m->nabr[j * NE + i] - 1, ilake = m->lakenabr[j * NE + i] - 1;
            const double B = m->edge[j * NE + i];
            double Q, nsf, dh, Ymean, s, CrossA;
            if (ilake >= 0) {
                nsf = yLakeStg[ilake];
                nsf = nsf < 0. ? 0. : nsf;
                Q = WeirFlow_jtoi(m->lake_zmin[ilake], nsf, m->z_surf[i], isf, m->z_surf[i], 0.6, B, 0.01);
            } else if (inabr >= 0) {
                nsf = uYsf[inabr];
                nsf = nsf < 0. ? 0. : nsf;
                dh = (isf + m->z_surf[i]) - (nsf + m->z_surf[inabr]);
                Ymean = avgY_sf(m->z_surf[i], isf, m->z_surf[inabr], nsf, m->depression[i]);
                Ymean = MIN(Ymean, MAXYSURF);
                if (Ymean <= 0.) {
                    Q = 0.;
                } else {
                    s = dh / m->Dist2Nabor[j * NE + i];
                    CrossA = Ymean * B;
                    if (s > 0 && isf <= 0) {
                        Q = 0.;
                    } else if (s < 0 && nsf <= 0) {
                        Q = 0.;
                    } else {
                        Q = ManningEquation(CrossA, m->avgRough[j * NE + i], Ymean, s);
                    }
                }
            } else {
                Q = 0;
                if (!m->close_boundary) {
                    if (isf > m->depression[i]) {
                        s = isf / m->Dist2Edge[j * NE + i] * 0.5;
                        if (s > 0.) {
                            Q = sqrt(s) * cbrt(isf * isf * isf * isf * isf) * B / m->Rough[i];
                        }
                    }
                }
            }
            QeleSurf[j * NE + i] = Q;
        }
        for (int j = 0; j < 3; j++) {
            const int inabr = m->nabr[j * NE + i] - 1, ilake = m->lakenabr[j * NE + i] - 1;
            double Q, dh, Ymean, grad, Kmean;
            if (ilake >= 0) {
                const double zl = m->lake_bathy_yi[m->lake_bathy_ptr[ilake]];
                dh = (uYgw[i] + m->z_bottom[i]) - (yLakeStg[ilake] + zl);
                if (dh > 0. && uYgw[i] <= 0.02) {
                    Q = 0.;
                } else if (dh < 0. && yLakeStg[ilake] <= 0.02) {
                    Q = 0.;
                } else {
                    Ymean = avgY_gw(uYgw[i], yLakeStg[ilake]);
                    grad = dh / m->Dist2Nabor[j * NE + i];
                    Kmean = 0.5 * (c.effKH[i] + c.effKH[inabr]);
                    Q = Kmean * grad * Ymean * m->edge[j * NE + i];
                }
            } else if (inabr >= 0) {
                dh = (uYgw[i] + m->z_bottom[i]) - (uYgw[inabr] + m->z_bottom[inabr]);
                if (dh > 0. && uYgw[i] <= 0.02) {
                    Q = 0.;
                } else if (dh < 0. && uYgw[inabr] <= 0.02) {
                    Q = 0.;
                } else {
                    Ymean = avgY_gw(uYgw[i], uYgw[inabr]);
                    grad = dh / m->Dist2Nabor[j * NE + i];
                    Kmean = 0.5 * (c.effKH[i] + c.effKH[inabr]);
                    Q = Kmean * grad * Ymean * m->edge[j * NE + i];
                }
            } else {
                Q = 0;
                if (!m->close_boundary) {
                    if (uYgw[i] > m->depression[i] * 10.) {
                        grad = uYgw[i] / m->Dist2Edge[j * NE + i] * 0.5;
                        if (grad > 0.) {
                            Q = c.effKH[i] * grad;
                        }
                    }
                }
            }
            QeleSub[j * NE + i] = Q * F->fu_Sub[i];
        }
    }
    PROF("loopB");
    /* bank-edge accumulations of LOOP B, ascending (cell, edge) order (MD_ElementFlux.cpp:52,121).
     * QLakeSub accumulates Q BEFORE the fu_Sub factor (line 121 precedes line 153). */
    if (m->lakeon && Nl > 0)
        for (int i = 0; i < Ne; i++) {
            if (m->iLake[i] > 0) continue;
            for (int j = 0; j < 3; j++) {
                const int ilake = m->lakenabr[j * NE + i] - 1;
                if (ilake >= 0) QLakeSurf[ilake] += QeleSurf[j * NE + i];
            }
            for (int j = 0; j < 3; j++) {
                const int ilake = m->lakenabr[j * NE + i] - 1;
                if (ilake >= 0) {
                    /* recover the un-scaled Q exactly: recompute it (fu_Sub may be != 1) */
                    const int inabr = m->nabr[j * NE + i] - 1;
                    const double zl = m->lake_bathy_yi[m->lake_bathy_ptr[ilake]];
                    double Q, dh = (uYgw[i] + m->z_bottom[i]) - (yLakeStg[ilake] + zl);
                    if (dh > 0. && uYgw[i] <= 0.02) Q = 0.;
                    else if (dh < 0. && yLakeStg[ilake] <= 0.02) Q = 0.;
                    else {
                        double Ymean = avgY_gw(uYgw[i], yLakeStg[ilake]);
                        double grad = dh / m->Dist2Nabor[j * NE + i];
                        double Kmean = 0.5 * (c.effKH[i] + c.effKH[inabr]);
                        Q = Kmean * grad * Ymean * m->edge[j * NE + i];
                    }
                    QLakeSub[ilake] += Q;
                }
            }
        }

    PROF("loopC_pre");
    /* ---- LOOP C (MD_f.cpp:37-40): fun_Seg_surface / fun_Seg_sub, MD_RiverFlux.cpp:100-126 ---- */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) schedule(static)
    for (int s = 0; s < Ns; s++) {
        const int ie = m->seg_iEle[s] - 1, ir = m->seg_iRiv[s] - 1;
        double isf = uYsf[ie] - qInfil[ie] + qExfil[ie];
        isf = MAX(0., isf);
        QsegSurf[s] = WeirFlow_jtoi(m->z_surf[ie], isf, m->z_surf[ie] - m->riv_depth[ir], uYriv[ir],
                                    m->z_surf[ie] + m->riv_zbank[ir], m->seg_Cwr[s], m->seg_length[s], m->depression[ie]);
        QsegSub[s] = flux_R2E_GW(uYriv[ir], m->z_surf[ie] - m->riv_depth[ir], uYgw[ie], m->z_bottom[ie], c.effKH[ie],
                                 m->riv_KsatH[ir], m->seg_length[s], m->riv_BedThick[ir]);
        QsegSub[s] *= F->fu_Sub[ie];
    }
    PROF("loopC");
    /* ---- LOOP D (MD_f.cpp:41-43): Flux_RiverDown, MD_RiverFlux.cpp:5-63 ---- */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) reduction(max : err) schedule(static)
    for (int i = 0; i < Nr; i++) {
        double Distance, A, Perem, R, s, n, sMean;
        const int iDown = m->riv_down[i] - 1;
        n = m->riv_avgRough[i];
        if (m->riv_toLake[i] >= 0) {
            Perem = CSperem[i];
            s = m->riv_BedSlope[i] + uYriv[i] * 2. / m->riv_Length[i];
            A = CSarea[i];
            R = (Perem <= 0.) ? 0. : (A / Perem);
            QrivDown[i] = ManningEquation(A, n, R, s);
        } else if (iDown >= 0) {
            sMean = (m->riv_BedSlope[i] + m->riv_BedSlope[iDown]) * 0.5;
            Distance = m->riv_Dist2DownStream[i];
            s = ((uYriv[i] - m->riv_depth[i]) - (uYriv[iDown] - m->riv_depth[iDown])) / Distance + sMean;
            A = CSarea[i];
            Perem = CSperem[i];
            R = (Perem <= ZERO) ? 0. : (A / Perem);
            QrivDown[i] = ManningEquation(A, n, R, s);
        } else {
            switch (m->riv_down[i]) {
                case -1:
                case -2:
                case -3:
                    Perem = CSperem[i];
                    s = m->riv_BedSlope[i] + uYriv[i] * 2. / m->riv_Length[i];
                    A = CSarea[i];
                    R = (Perem <= 0.) ? 0. : (A / Perem);
                    QrivDown[i] = ManningEquation(A, n, R, s);
                    break;
                case -4:
                    QrivDown[i] = CSarea[i] * sqrt(GRAV * uYriv[i]) * 60.;
                    break;
                default:
                    QrivDown[i] = 0.;
                    if (SHUD_ERRRIVBC > err) err = SHUD_ERRRIVBC;
            }
        }
    }
    for (int i = 0; i < Nr; i++)
        if (m->riv_toLake[i] >= 0) QLakeRivIn[m->riv_toLake[i]] += QrivDown[i]; /* MD_RiverFlux.cpp:24 */
    /* ---- LOOP E (MD_f.cpp:44-47) ---- */
    for (int l = 0; l < Nl; l++) {
        qLakeEvap[l] = MIN(qLakeEvap[l], qLakePrcp[l] + yLakeStg[l]);
        qLakeEvap[l] = MAX(0, qLakeEvap[l]);
    }
    PROF("loopD");
    /* ---- PassValue, MD_f.cpp:217-240: ordered scatter-add, ascending source index ---- */
    for (int s = 0; s < Ns; s++) {
        const int ie = m->seg_iEle[s] - 1, ir = m->seg_iRiv[s] - 1;
        QrivSurf[ir] += QsegSurf[s];
        QrivSub[ir] += QsegSub[s];
        Qe2rS[ie] += -QsegSurf[s];
        Qe2rG[ie] += -QsegSub[s];
    }
    for (int i = 0; i < Nr; i++) {
        const int iDown = m->riv_down[i] - 1;
        if (iDown >= 0 && m->riv_toLake[i] <= 0) QrivUp[iDown] += -QrivDown[i];
    }

    PROF("PassValue");
    /* ================= f_applyDY, src/ModelData/MD_f.cpp:52-191 ================= */
#pragma omp parallel for num_threads(nthreads) if (nthreads > 1) reduction(max : err) schedule(static)
    for (int i = 0; i < Ne; i++) {
        const double area = m->area[i];
        double SurfTot = Qe2rS[i], SubTot = Qe2rG[i];
        int e = 0;
        for (int j = 0; j < 3; j++) {
            SurfTot += QeleSurf[j * NE + i];
            SubTot += QeleSub[j * NE + i];
            if (bad_nan(QeleSurf[j * NE + i]) || bad_nan(QeleSub[j * NE + i])) e = SHUD_ERRNAN;
        }
        double dsf = F->qEleNetPrep[i] - qInfil[i] + qExfil[i] - SurfTot / area - qEs[i];
        double dus = qInfil[i] - qRech[i] - qEu[i] - qTu[i];
        double dgw = qRech[i] - qExfil[i] - SubTot / area - qEg[i] - qTg[i];
        if (m->iBC[i] > 0) {
            dgw = 0;
        } else if (m->iBC[i] < 0) {
            dgw += QBC[i] / area;
        }
        if (m->iSS[i] > 0) {
            dsf += m->QSS[i] / area;
        } else if (m->iSS[i] < 0) {
            dgw += m->QSS[i] / area;
        }
        dus /= m->Sy[i];
        dgw /= m->Sy[i];
        if (m->iLake[i] > 0) {
            dsf = 0.; dus = 0.; dgw = 0.;
        }
        DY[i] = dsf; DY[i + NE] = dus; DY[i + 2 * NE] = dgw;
        STORE(QeleSurfTot, i, SurfTot); STORE(QeleSubTot, i, SubTot);
        if (e > err) err = e;
    }
    for (int i = 0; i < Nr; i++) {
        double d;
        if (m->riv_BC[i] > 0) {
            d = 0.;
        } else {
            d = (-QrivUp[i] - QrivSurf[i] - QrivSub[i] - QrivDown[i] + qBC[i]) / m->riv_Length[i];
            if (d < -1. * CSarea[i]) d = -1. * CSarea[i];
            d = fun_dAtodY(d, topWidth[i], m->riv_bankslope[i]);
        }
        DY[3 * NE + i] = d;
    }
    for (int l = 0; l < Nl; l++) {
        DY[3 * NE + Nr + l] = qLakePrcp[l] - qLakeEvap[l] +
                              (QLakeRivIn[l] - QLakeRivOut[l] + QLakeSub[l] + QLakeSurf[l]) / y2LakeArea[l];
    }

    PROF("applyDY");
    /* ---- diagnostics ---- */
    if (diag) {
#define COPY(dst, src, n) do { if (diag->dst) memcpy(diag->dst, src, sizeof(double) * (size_t)(n)); } while (0)
        COPY(qEleInfil, qInfil, Ne); COPY(qEleExfil, qExfil, Ne); COPY(qEleRecharge, qRech, Ne);
        COPY(qEs, qEs, Ne); COPY(qEu, qEu, Ne); COPY(qEg, qEg, Ne); COPY(qTu, qTu, Ne); COPY(qTg, qTg, Ne);
        COPY(u_effKH, c.effKH, Ne); COPY(u_satn, c.satn, Ne);
        COPY(QeleSurf, QeleSurf, 3 * NE); COPY(QeleSub, QeleSub, 3 * NE);
        COPY(Qe2r_Surf, Qe2rS, Ne); COPY(Qe2r_Sub, Qe2rG, Ne);
        COPY(QsegSurf, QsegSurf, Ns); COPY(QsegSub, QsegSub, Ns);
        COPY(QrivSurf, QrivSurf, Nr); COPY(QrivSub, QrivSub, Nr); COPY(QrivUp, QrivUp, Nr); COPY(QrivDown, QrivDown, Nr);
        COPY(y2LakeArea, y2LakeArea, Nl); COPY(QLakeSurf, QLakeSurf, Nl); COPY(QLakeSub, QLakeSub, Nl);
        COPY(QLakeRivIn, QLakeRivIn, Nl); COPY(QLakeRivOut, QLakeRivOut, Nl);
        COPY(qLakeEvap, qLakeEvap, Nl); COPY(qLakePrcp, qLakePrcp, Nl);
    }
    return err;
}


/* =====================================================================================================
 * Land-surface step: restatement of Model_Data::tReadForcing
 * (src/ModelData/MD_ET.cpp:21-281) and Model_Data::ET (MD_ET.cpp:282-342) with the helpers of
 * src/Equations/is_sm_et.hpp / is_sm_et.cpp, Equations.hpp:66-72 and functions.hpp:191-201.
 * Pinned by tests/golden/ccw.land.npz: 30 consecutive steps of the reference itself (oracle/ref_driver.cpp
 * --land-seq), rain, snow accumulation and melt, interception, day and night, terrain radiation on.
 * ===================================================================================================== */
#define L_SecADay 86400          /* Macros.hpp:43 */
#define L_dTdZ 0.0065            /* Macros.hpp:50 */
#define L_Tsnow (-3.0)           /* Macros.hpp:59 */
#define L_Train 1.0              /* Macros.hpp:60 */
#define L_To 0.0                 /* Macros.hpp:61 */
#define L_ROUGHNESS_WATER 0.00137 /* Macros.hpp:62 */
#define L_CONST_RH 0.01          /* Macros.hpp:63 */
#define L_IC_MAX 0.0002          /* Macros.hpp:65 */
#define L_VON_KARMAN 0.4         /* Macros.hpp:70 */
#define L_Cp 1.013e-3            /* Macros.hpp:72 */
#define L_NA (-9999)             /* Macros.hpp:83 */
static double lmin(double a, double b) { return a > b ? b : a; } /* functions.hpp:117-123 */
static double lmax(double a, double b) { return a < b ? b : a; }
static double frozen_fraction(double T, double high, double low) { /* functions.hpp:191-201 */
    if (T > high) return 0;
    if (T < low) return 1;
    return lmin(1.0, lmax((high - T) / (high - low), 0.0));
}

long shud_oracle_cryo_size(const shud_mesh *m, const shud_land *L) {
    return 8 + (long)m->Ne * (3 + (long)L->FT_surf_day + (long)L->FT_sub_day);
}

int shud_oracle_land_step(const shud_mesh *m, const shud_land *L, const shud_land_step *S, double *yEleSnow,
                          double *yEleIS, const shud_land_out *out, double *cryo) {
    const int Ne = m->Ne;
    int rc = 0;
    const double DT_min = S->dt_min;
    /* _AccTemp::push(x, tnow), AccTemperature.hpp:47-57: both accumulators see the same pushes, so the day sum,
     * its count and the day clock are shared; a daily mean is queued when >= 1440 min have passed */
    const int Ls = (int)L->FT_surf_day, Lb = (int)L->FT_sub_day;
    double *Tacc = NULL, *ACCs = NULL, *ACCb = NULL, *ringS = NULL, *ringB = NULL;
    int do_push = 0, popS = 0, popB = 0, slotS = 0, slotB = 0;
    double nday = 0., sizeS = 1., sizeB = 1.;
    if (L->cryosphere) {
        if (!cryo) return -2;
        Tacc = cryo + 8; ACCs = Tacc + Ne; ACCb = ACCs + Ne; ringS = ACCb + Ne; ringB = ringS + (size_t)Ls * Ne;
        cryo[1] += 1.;                     /* N_of_day++ */
        nday = cryo[1];
        do_push = (S->t - cryo[0]) >= 1440.;
        if (do_push) {
            /* que.push; if (size > MaxLen) pop the oldest: ring of MaxLen slots, the new value takes the slot
             * of the value it displaces */
            int size = (int)cryo[2], head = (int)cryo[3];
            if (size == Ls) { popS = 1; slotS = head; cryo[3] = (head + 1) % Ls; }
            else { slotS = (head + size) % Ls; cryo[2] = size + 1; }
            size = (int)cryo[4]; head = (int)cryo[5];
            if (size == Lb) { popB = 1; slotB = head; cryo[5] = (head + 1) % Lb; }
            else { slotB = (head + size) % Lb; cryo[4] = size + 1; }
            cryo[0] = S->t;
            cryo[1] = 0.;
        }
        sizeS = cryo[2]; sizeB = cryo[4];
    }
    for (int i = 0; i < Ne; i++) {
        /* ---------------- tReadForcing, MD_ET.cpp:21-281 ---------------- */
        const int idx = L->iForc[i] - 1;
        const double *row = S->forc + 5 * idx;
        double t_prcp = row[0] * L->cPrep;
        const double t0 = row[1];
        const double Zt = L->forc_z[idx], Zi = m->z_surf[i];
        double t_temp; /* TemperatureOnElevation, Equations.hpp:66-72 */
        if (fabs(Zi - L_NA) < ZERO || fabs(Zt - L_NA) < ZERO) t_temp = t0;
        else t_temp = t0 + (Zt - Zi) * L_dTdZ;
        t_temp = t_temp + L->cTemp;
        const double lai = S->lai[L->iLC[i] - 1] * L->cLAItsd;
        const double t_mf = S->mf[L->iMF[i] - 1] * L->cMF / 1440.;
        const double dswrf_h = row[4];
        double dswrf_t = dswrf_h, factor = 1.0;
        if (L->terrain_radiation) {
            if (S->tsr_n < 0) {
                factor = 0.0;
            } else {
                double num = 0.0;
                const double cap = L->rad_factor_cap, cosz_min = L->rad_cosz_min;
                if (S->tsr_den > 0.0 && S->tsr_n > 0) {
                    const double nx = L->nx[i], ny = L->ny[i], nz = L->nz[i];
                    for (int k = 0; k < S->tsr_n; k++) {
                        const double wdt = S->tsr_wdt[k];
                        if (!(wdt > 0.0)) continue;
                        const double sx = S->tsr_sx[k], sy = S->tsr_sy[k], sz = S->tsr_sz[k];
                        const double cosi = nx * sx + ny * sy + nz * sz;
                        if (!(cosi > 0.0) || !isfinite(cosi)) continue;
                        double denom = sz;
                        if (denom < cosz_min) denom = cosz_min;
                        if (!(denom > 0.0) || !isfinite(denom)) continue;
                        double fk = cosi / denom;
                        if (!isfinite(fk) || !(fk > 0.0)) continue;
                        if (fk > cap) fk = cap;
                        num += wdt * fk;
                    }
                }
                double feff = 0.0;
                if (S->tsr_den > 0.0) {
                    feff = num / S->tsr_den;
                    if (!isfinite(feff) || !(feff > 0.0)) feff = 0.0;
                    if (feff > L->rad_factor_cap) feff = L->rad_factor_cap;
                }
                factor = feff;
            }
            dswrf_t = dswrf_h * factor;
        }
        double t_rn = L->radiation_is_net ? dswrf_t : dswrf_t * (1 - L->Albedo[i]);
        const double Uz = fabs(row[3]) + 0.001;
        double t_rh = row[2];
        t_prcp = t_prcp * 0.001 / 1440.;
        t_rn = t_rn * 1.0e-6;
        t_rh = lmin(lmax(t_rh, L_CONST_RH), 1.0);
        const double lambda = 2.501 - 0.002361 * t_temp;                       /* LatentHeat */
        const double Gamma = 0.0016286 * L->FixPressure[i] / lambda;           /* PsychrometricConstant */
        const double es = 0.6108 * exp(17.27 * t_temp / (t_temp + 237.3));     /* VaporPressure_Sat */
        const double ea = es * t_rh;
        const double ed = es - ea;
        const double tt = t_temp + 237.3;
        const double Delta = 4098. * es / (tt * tt);                           /* SlopeSatVaporPressure */
        const double rho = 3.486 * L->FixPressure[i] / (275. + t_temp);        /* AirDensity */
        const int lake = m->iLake[i] > 0;
        double GroundHeatFlux;
        if (lake) GroundHeatFlux = 0.;
        else if (lai > 0) GroundHeatFlux = 0.4 * exp(-0.5 * lai) * t_rn;
        else GroundHeatFlux = 0.1 * t_rn;
        const double RG = t_rn - GroundHeatFlux;
        /* WindProfile(2.0, Uz, windH, 0., ROUGHNESS_WATER) */
        const double U2 = Uz * log((2.0 - 0.) / L_ROUGHNESS_WATER) / log((L->windH[i] - 0.) / L_ROUGHNESS_WATER);
        double pm_ow; /* PET_PM_openwater, is_sm_et.cpp:57-64 */
        {
            double ETp = (Delta * RG * L_SecADay + Gamma * 6.43 * (1.0 + 0.536 * U2) * ed) / (Delta + Gamma);
            ETp = ETp / lambda;
            ETp = ETp * 0.001 / L_SecADay;
            pm_ow = ETp;
        }
        const double qPotEvap = L->cETP * pm_ow * 60.;
        double qPotTran, etp;
        if (lake) {
            qPotTran = L->cETP * 0.;
            etp = qPotEvap;
        } else if (lai <= 0.) {
            qPotTran = L->cETP * 0.;
            etp = qPotEvap;
        } else {
            const double hc = lai * 0.5;
            const double Zmeasure = hc * 1.3333;
            double ra; /* AerodynamicResistance(Uz, hc, Zmeasure, Zmeasure), is_sm_et.hpp */
            {
                const double d = 0.67 * hc, Z_om = 0.123 * hc, Z_ov = 0.0123 * hc;
                ra = log(fabs(Zmeasure - d) / Z_om) * log(fabs(Zmeasure - d) / (Z_ov)) / (L_VON_KARMAN * L_VON_KARMAN * Uz);
            }
            if (ra <= 0.0 || isnan(ra) || isinf(ra) || fabs(ra - L_NA) < ZERO) rc = 10; /* CheckNonZero */
            const double rs = 200. / lai; /* BulkSurfaceResistance(lai) */
            double pm; /* PET_Penman_Monteith, is_sm_et.cpp:32-56 */
            {
                const double E_rad = Delta * RG;
                const double E_air = rho * L_Cp * ed / ra;
                const double r_sa = rs / ra;
                double ETp = (E_rad + E_air) / (Delta + Gamma * (1 + r_sa));
                ETp = ETp / lambda;
                ETp = ETp * 0.001;
                pm = ETp;
            }
            qPotTran = L->cETP * pm * 60.;
            etp = qPotTran * m->VegFrac[i] + qPotEvap * (1. - m->VegFrac[i]);
            if (isnan(qPotTran)) rc = 10;
        }
        /* ---------------- ET, MD_ET.cpp:282-342 ---------------- */
        const double T = t_temp, prcp = t_prcp, MF = t_mf;
        double snStg = yEleSnow[i];
        const double snFrac = frozen_fraction(T, L_Train, L_Tsnow);
        double fu_Sub = 1., fu_Surf = 1.;
        if (L->cryosphere) { /* MD_ET.cpp:301-307 */
            Tacc[i] += T;
            if (do_push) {
                const double x = Tacc[i] / nday;
                ACCs[i] += x;
                if (popS) ACCs[i] -= ringS[(size_t)slotS * Ne + i];
                ringS[(size_t)slotS * Ne + i] = x;
                ACCb[i] += x;
                if (popB) ACCb[i] -= ringB[(size_t)slotB * Ne + i];
                ringB[(size_t)slotB * Ne + i] = x;
                Tacc[i] = 0.;
            }
            const double ta_surf = ACCs[i] / sizeS, ta_sub = ACCb[i] / sizeB;
            fu_Sub = 1. - frozen_fraction(ta_sub, L->FT_sub_max, L->FT_sub_min);
            fu_Surf = 1. - frozen_fraction(ta_surf, L->FT_surf_max, L->FT_surf_min);
        }
        const double snAcc = snFrac * prcp;
        double snMelt = (T > L_To ? (T - L_To) * MF : 0.);
        snMelt = lmin(lmax(0., snStg / DT_min), lmax(0., snMelt));
        snStg += (snAcc - snMelt) * DT_min;
        const double vgFrac = m->VegFrac[i];
        double icStg = (vgFrac > ZERO) ? (yEleIS[i] / vgFrac) : 0.0;
        double icAcc, icEvap;
        if (lai > ZERO) {
            const double icMax = L->cISmax * L_IC_MAX * lai;
            icAcc = lmin(prcp - snAcc, lmax(0., (icMax - icStg) / DT_min));
            icEvap = lmin(lmax(0., icStg / DT_min), qPotEvap);
        } else {
            icAcc = 0.;
            icEvap = 0.;
        }
        icStg += (icAcc - icEvap) * DT_min;
        yEleIS[i] = icStg * vgFrac;
        yEleSnow[i] = snStg;
        if (out) {
            if (out->qElePrep) out->qElePrep[i] = t_prcp;
            if (out->qPotEvap) out->qPotEvap[i] = qPotEvap;
            if (out->qPotTran) out->qPotTran[i] = qPotTran;
            if (out->qEleETP) out->qEleETP[i] = etp;
            if (out->t_lai) out->t_lai[i] = lai;
            if (out->t_temp) out->t_temp[i] = t_temp;
            if (out->t_mf) out->t_mf[i] = t_mf;
            if (out->qEleNetPrep) out->qEleNetPrep[i] = (1. - snFrac) * prcp + snMelt - icAcc * vgFrac;
            if (out->qEleE_IC) out->qEleE_IC[i] = icEvap * vgFrac;
            if (out->fu_Surf) out->fu_Surf[i] = fu_Surf;
            if (out->fu_Sub) out->fu_Sub[i] = fu_Sub;
            if (out->rn_factor) out->rn_factor[i] = factor;
            if (out->yEleSnow) out->yEleSnow[i] = snStg;
            if (out->yEleIS) out->yEleIS[i] = icStg * vgFrac;
        }
    }
    return rc;
}
