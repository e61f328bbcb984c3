// shud_phys.cuh - the constitutive relations of the SHUD right-hand side as FP64 device
// functions.  Written from the reference's behaviour; each function names the reference
// source it has to agree with (paths relative to the reference root).  Compiled with
// -fmad=false so that, like the reference's x86-64 build, no product-sum is contracted.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace shud {

// literal constants of src/Model/Macros.hpp:31-35,46,51,67 (kept as written there)
constexpr double kEPSILON = 0.005;
constexpr double kZERO = 1.0e-10;
constexpr double kEPS_SLOPE = 0.05e-6;
constexpr double kPI = 3.1415926;
constexpr double kGRAV = 9.8;
constexpr double kMAXYSURF = 0.5;
constexpr double kNA_VALUE = -9999.0;
constexpr double kFieldCapacityRatio = 0.75;

// src/Equations/functions.hpp:117-123: (a>b?b:a) / (a<b?b:a) - not fmin/fmax
__device__ __forceinline__ double dmin(double a, double b) { return a > b ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return a < b ? b : a; }

// CheckNANi (src/Equations/functions.cpp:90-96) and CheckNonNegative (:148-154)
__device__ __forceinline__ bool not_finite(double x) { return !(fabs(x) <= 1.79769313486231570815e308); }
__device__ __forceinline__ bool bad_nonneg(double x) {
    return x < 0.0 || not_finite(x) || fabs(x - kNA_VALUE) < kZERO;
}

// Square root.  Default: the IEEE library sqrt (a 49-instruction CALL).  -DSHUD_LEAN_SQRT swaps in an inline
// MUFU.RSQ64H-seeded Newton sequence (<= 1 ulp); measured on B200 it does not change the kernel time
// (137.4 vs 137.2 us: the kernel is latency-, not instruction-bound there), so it stays off.
__device__ __forceinline__ double fsqrt(double x) {
#ifdef SHUD_LEAN_SQRT
    if (!(x > 2.3e-308 && x < 1.7e308)) return sqrt(x);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-(x * y), y, 1.0);
    y = fma(y * 0.5, e, y);
    e = fma(-(x * y), y, 1.0);
    y = fma(y * 0.5, e, y);
    double r = x * y;
    const double d = fma(-r, r, x);
    return fma(d, 0.5 * y, r);
#else
    return sqrt(x);
#endif
}

// Division by a STATIC quantity.  Default: the IEEE division the reference performs.  With SHUD_RCP the
// host stores the reciprocal of that static array instead and the kernel multiplies (<= 1.5 ulp away
// from the quotient; inside the 1e-12 parity tolerance, not bit-identical).
#ifdef SHUD_RCP
#define SHUD_DIVS(x, d_or_rcp) ((x) * (d_or_rcp))
#else
#define SHUD_DIVS(x, d_or_rcp) ((x) / (d_or_rcp))
#endif

// ManningEquation, src/Equations/Equations.hpp:54-63 (pow23 = cbrt^2, :36-39)
__device__ __forceinline__ double manning(double area, double rough, double R, double S) {
    const double c = cbrt(R);
    const double p23 = c * c;
    if (S > 0) return fsqrt(S) * area * p23 / rough;
    return -1.0 * fsqrt(-S) * area * p23 / rough;
}
// same, for a cell edge: `rough` is the static avgRough[j] (or its reciprocal under SHUD_RCP)
__device__ __forceinline__ double manning_edge(double area, double rough, double R, double S) {
    const double c = cbrt(R);
    const double p23 = c * c;
    if (S > 0) return SHUD_DIVS(fsqrt(S) * area * p23, rough);
    return SHUD_DIVS(-1.0 * fsqrt(-S) * area * p23, rough);
}

// effKH, src/Equations/Equations.cpp:116-134.  Range violation -> *err = 13 (myexit(ERRDATAIN)).
__device__ __forceinline__ double eff_kh(double ygw, double aqd, double macD, double kmac, double af, double kmx,
                                         int *err) {
    double k;
    if (macD <= kZERO || ygw < aqd - macD) {
        k = kmx;
    } else if (ygw > aqd) {
        k = (kmac * macD * af + kmx * (aqd - macD * af)) / aqd;
    } else {
        const double h = ygw - (aqd - macD);
        k = (kmac * h * af + kmx * (aqd - macD + h * (1 - af))) / ygw;
    }
    if (k < 0. || k > 1e9) *err = 13;
    return k;
}

// van Genuchten-Mualem relative conductivity, satKfun, src/Equations/Equations.cpp:136-141
// x^a for 0 < x < 1.  SHUD_POW_EXPLOG: exp(a log x) - |a log x| ulp-level error (<= ~2e-14 relative for the
// exponents here) instead of pow()'s <= 2 ulp, at well under half the instructions.
__device__ __forceinline__ double pow01(double x, double a) {
#ifdef SHUD_POW_EXPLOG
    return exp(a * log(x));
#else
    return pow(x, a);
#endif
}
__device__ __forceinline__ double sat_kr(double s, double n) {
    // -1 + (1 - s^(n/(n-1)))^((n-1)/n): the two pow() as two trips of one loop, so that the kernel holds one copy
    // of pow()'s ~150 instructions instead of two (instruction-cache footprint of the cell kernel)
    double x = s, e = n / (n - 1.), r = 0.;
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        r = pow01(x, e);
        x = 1. - r;
        e = (n - 1.) / n;
    }
    const double t = -1. + r;
    return fsqrt(s) * t * t;
}

// SoilMoistureStress, src/Equations/is_sm_et.cpp:131-142
__device__ __forceinline__ double soil_moisture_stress(double thetaS, double thetaR, double satn) {
    const double fc = thetaS * kFieldCapacityRatio;
    double b = (satn * (thetaS - thetaR) - thetaR) / (fc - thetaR);
    b = dmin(dmax(0., b), 1.);
    return 0.5 * (1 - cos(kPI * b));
}

// Model_Data::WeirFlow_jtoi, src/ModelData/MD_RiverFlux.cpp:65-98 (positive = j -> i)
__device__ __forceinline__ double weir_jtoi(double zi, double yi, double zj, double yj, double zbank, double cwr,
                                            double width, double thr) {
    const double hi = yi + zi, hj = yj + zj, dh = hj - hi;
    double y = hi - zbank, Q = 0.;
    if (dh > 0.) {
        if (y > 0. && yj > thr) {
            if (hi > zbank) y = dh;
            Q = cwr * fsqrt(2. * kGRAV * y) * width * y * 60.;
        }
    } else {
        if (y > 0. && yi > thr) {
            if (hj > zbank) y = -dh;
            Q = -1. * cwr * fsqrt(2. * kGRAV * y) * width * y * 60.;
        }
    }
    return Q;
}

// flux_R2E_GW, src/Equations/Flux_RiverElement.cpp:11-55
__device__ __forceinline__ double flux_r2e_gw(double yr, double zr, double ye, double ze, double kele, double kriv,
                                              double L, double bed) {
    if (kele < kZERO || kriv < kZERO) return 0.;
    const double K = (kele * 1. + kriv * 1.) / (1. + 1.);  // meanArithmetic(.,.,1,1), Equations.hpp:50-52
    const double he = ye + ze, hr = yr + zr, dh = hr - he;
    double Q = 0.;
    if (dh > kZERO) {
        const double A = (he > zr) ? (yr + (he - zr)) * .5 * L : yr * L;
        if (!(yr < kEPSILON)) Q = A * K * (dh / bed);
    } else if (dh < -kZERO) {
        if (ye > kZERO) {
            const double A = (yr + (he - zr)) * .5 * L;
            Q = A * K * (dh / bed);
        }
    }
    return Q;
}

// fun_dAtodY + Quadratic, src/Equations/functions.hpp:125-153
__device__ __forceinline__ double dA_to_dY(double dA, double wtop, double s) {
    if (dA == 0.) return 0.;
    if (fabs(s) < kEPS_SLOPE) return dA / wtop;
    s = fabs(s);
    const double cc = wtop * wtop + 4 * s * dA;
    if (cc < kZERO) return -1. * wtop / (2. * s);
    return (-wtop + sqrt(cc)) / (2 * s);
}

// LakeBathymetry::toparea, src/classes/Lake.cpp:59-78 (slope over yi[i]-y, as the reference has it)
__device__ __forceinline__ double lake_toparea(const double *yi, const double *ai, int n, double y) {
    double ta = ai[0];
    if (y <= yi[0]) return ta;
    for (int i = 1; i < n; i++) {
        if (y < yi[i]) {
            const double da = ai[i] - ta, dy = yi[i] - y;
            return da / dy * (y - yi[i - 1]) + ta;
        }
        ta = ai[i];
    }
    return ta;
}

// trapezoid cross-section, _River::updateRiver, src/classes/River.cpp:49-62 + River.hpp:115-128
struct RivGeom {
    double topWidth, csArea, csPerem;
};
__device__ __forceinline__ RivGeom riv_geom(double y, double w0, double s) {
    RivGeom g;
    const double ys = y * s;
    g.topWidth = dmax(ys * 2.0 + w0, 0.);  // fixMaxValue(x,0): x<0 ? 0 : x
    g.csArea = dmax(y * (w0 + ys), 0.);
    g.csPerem = dmax(2.0 * sqrt(y * y + ys * ys) + w0, 0.);
    return g;
}

// Flux_RiverDown, src/ModelData/MD_RiverFlux.cpp:5-63.
//   yraw  : Y[iRIV] (geometry is updated from it BEFORE a stage BC overrides uYriv, MD_update.cpp:146-158)
//   ystg  : uYriv[i] (after the BC override);  ystg_dn/depth_dn/slope_dn : downstream reach (if down>0)
__device__ __forceinline__ double river_down(double yraw, double ystg, double w0, double bankslope, double length,
                                             double bedslope, double depth, double rough, double dist, int down,
                                             int toLake, double ystg_dn, double depth_dn, double slope_dn, int *err) {
    const RivGeom g = riv_geom(yraw, w0, bankslope);
    if (toLake >= 0 || (down <= 0 && down >= -3)) {
        // into a lake, or outlet types -1/-2/-3: zero-depth-gradient
        if (toLake < 0 && down == 0) { *err = 1; return 0.; }
        const double s = bedslope + ystg * 2. / length;
        const double R = (g.csPerem <= 0.) ? 0. : (g.csArea / g.csPerem);
        return manning(g.csArea, rough, R, s);
    }
    if (down > 0) {
        const double sMean = (bedslope + slope_dn) * 0.5;
        const double s = ((ystg - depth) - (ystg_dn - depth_dn)) / dist + sMean;
        const double R = (g.csPerem <= kZERO) ? 0. : (g.csArea / g.csPerem);
        return manning(g.csArea, rough, R, s);
    }
    if (down == -4) return g.csArea * sqrt(kGRAV * ystg) * 60.;  // critical depth
    *err = 1;  // "River Routing Boundary Condition Type Is Wrong" -> exit(1)
    return 0.;
}

// ------------------------------------------------------------------------------------------
// per-cell vertical processes
// ------------------------------------------------------------------------------------------
struct CellParams {  // static, per cell
    double aqd, sy, infD, infKsatV, macKsatV, hAreaF, thetaS, thetaR, thetaFC, beta, ksatV;
    double vegFrac, impAF, wetland, rootReach;
};
struct CellForc {  // fixed between forcing steps
    double netPrep, potEvap, potTran, lai, fuSurf, fuSub;
};
struct CellVert {  // results
    double Es, Eu, Eg, Tu, Tg, eic, iBeta;
    double satn, infil, exfil, rech;
    int err;
};

// The cell's vertical processes in two independent halves (so that two warp roles can share them):
//   cell_et   : f_etFlux, src/ModelData/MD_ET.cpp:343-404 (uses the saturation of the PREVIOUS call)
//   cell_soil : _Element::updateElement (src/classes/Element.cpp:347-384, minus the dead u_phius / u_effkInfi
//               stores) -> Flux_Infiltration (Element.cpp:271-303) -> Flux_Recharge (Element.cpp:304-335), with the
//               fu_Surf / fu_Sub factors of MD_ElementFlux.cpp:24-34
__device__ __forceinline__ void cell_et(const CellParams &p, const CellForc &f, double ysf, double yus, double ygw,
                                        double satn_prev, double eic_in, CellVert &r) {
    {
        const double va = p.vegFrac, vb = 1. - p.vegFrac, pj = 1. - p.impAF;
        const double ib = soil_moisture_stress(p.thetaS, p.thetaR, satn_prev);
        double Es, Eu = 0., Eg = 0., Tu = 0., Tg = 0., eic = eic_in;
        Es = dmin(dmax(0., ysf), f.potEvap) * vb;
        if (Es < f.potEvap) {
            if (ygw > p.wetland) Eg = dmin(dmax(0., ygw), f.potEvap - Es) * pj * vb;
            else Eu = dmin(dmax(0., yus), ib * (f.potEvap - Es)) * pj * vb;
        }
        if (f.lai > kZERO) {
            if (eic >= f.potTran) {
                eic = f.potTran * pj * va;
            } else if (ygw > p.rootReach) {
                Tg = dmin(dmax(0., ygw), (f.potTran - eic)) * pj * va;
            } else {
                Tu = dmin(dmax(0., yus), ib * (f.potTran - eic)) * pj * va;
            }
        } else {
            eic = 0.;
        }
        const double trans = Tg + Tu, evapo = Eu + Eg + Es, eta = eic + evapo + trans;
        if (bad_nonneg(Es) || bad_nonneg(Eu) || bad_nonneg(Eg) || bad_nonneg(Tu) || bad_nonneg(Tg) ||
            not_finite(eta) || not_finite(evapo) || not_finite(trans))
            r.err = 10;
        r.Es = Es; r.Eu = Eu; r.Eg = Eg; r.Tu = Tu; r.Tg = Tg; r.eic = eic; r.iBeta = ib;
    }
}

// cell_soil in two steps, so that a caller short of registers can fetch the second step's inputs after the pow()s:
//   cell_soil_state : updateElement -> deficit, theta, satn, satKr     (inputs aqd, thetaS, thetaR, beta)
//   cell_soil_flux  : Flux_Infiltration + Flux_Recharge                 (everything else)
struct SoilState {
    double deficit, theta, satn, satKr;
};
__device__ __forceinline__ SoilState cell_soil_state(double aqd, double thetaS, double thetaR, double beta, double yus,
                                                     double ygw) {
    SoilState s;
    double deficit = aqd - ygw, satn, theta, satKr;
    if (deficit <= 0.) {
        deficit = 0.;
        satn = 1.;
        theta = thetaS;
    } else {
        theta = yus / deficit * thetaS;
        satn = (theta - thetaR) / (thetaS - thetaR);
    }
    if (satn > 0.99) {
        satn = 1.0; satKr = 1.0; theta = thetaS;
    } else if (satn <= kZERO) {
        satn = 0.; satKr = 0.; theta = thetaR;
    } else {
        satKr = sat_kr(satn, beta);
    }
    s.deficit = deficit; s.theta = theta; s.satn = satn; s.satKr = satKr;
    return s;
}
__device__ __forceinline__ void cell_soil_flux(const CellParams &p, const CellForc &f, double ysf, double yus, double ygw,
                                               const SoilState &st, CellVert &r) {
    const double deficit = st.deficit, theta = st.theta, satn = st.satn, satKr = st.satKr;
    const double kmax = p.infKsatV * (1. - p.hAreaF) + p.macKsatV * p.hAreaF;
    r.satn = satn;
    // ---- infiltration / exfiltration ----
    {
        const double av = ysf + f.netPrep;
        double qi = 0., qex = 0.;
        if (ygw + yus > p.aqd || deficit < yus) {
            qex = fabs(ygw + yus - p.aqd) / p.aqd * kmax;
        } else if (av > 0. && deficit > p.infD) {
            const double grad = 1. + av / p.infD;
            double keff;
            if (av > kmax) keff = p.infKsatV * (1 - p.hAreaF) + p.hAreaF * p.macKsatV * satn;
            else if (av > p.infKsatV) keff = satKr * p.infKsatV * (1 - p.hAreaF) + p.hAreaF * p.macKsatV * satn;
            else keff = satKr * p.infKsatV * (1 - p.hAreaF);
            qi = grad * keff;
            qi = dmin(av, dmax(0., qi));
        }
        r.infil = qi * f.fuSurf;
        r.exfil = qex * f.fuSurf;
    }
    // ---- recharge ----
    {
        double qr;
        if (ygw > p.aqd - p.infD && yus < deficit) {
            qr = 0.;
        } else {
            double grad = 0.;
            if (theta > p.thetaR && !(yus <= kEPSILON)) {
                grad = (theta - p.thetaR) / (p.thetaFC - p.thetaR);
                grad = dmax(grad, 0.);
            }
            if (p.infKsatV <= 0. || p.ksatV <= 0.) {
                qr = 0.;
            } else {
                const double ku = p.infKsatV * satKr;
                // meanHarmonic(ku, KsatV, deficit, Ygw), src/Equations/Equations.hpp:45-48
                const double ke = (ku * p.ksatV) * (deficit + ygw) / (deficit * p.ksatV + ygw * ku);
                qr = grad * ke;
            }
        }
        r.rech = qr * f.fuSub;
    }
}
__device__ __forceinline__ void cell_soil(const CellParams &p, const CellForc &f, double ysf, double yus, double ygw,
                                          CellVert &r) {
    const SoilState st = cell_soil_state(p.aqd, p.thetaS, p.thetaR, p.beta, yus, ygw);
    cell_soil_flux(p, f, ysf, yus, ygw, st, r);
}

__device__ __forceinline__ CellVert cell_vertical(const CellParams &p, const CellForc &f, double ysf, double yus,
                                                  double ygw, double satn_prev, double eic_in) {
    CellVert r;
    r.err = 0;
    cell_et(p, f, ysf, yus, ygw, satn_prev, eic_in, r);
    cell_soil(p, f, ysf, yus, ygw, r);
    return r;
}

// u_satn alone (priming the carried state), _Element::updateElement, src/classes/Element.cpp:349-368
__device__ __forceinline__ double cell_satn(double aqd, double thetaS, double thetaR, double yus, double ygw) {
    const double deficit = aqd - ygw;
    double satn;
    if (deficit <= 0.) satn = 1.;
    else satn = (yus / deficit * thetaS - thetaR) / (thetaS - thetaR);
    if (satn > 0.99) satn = 1.0;
    else if (satn <= kZERO) satn = 0.;
    return satn;
}

// overland flux through one edge to a neighbouring cell, fun_Ele_surface, src/ModelData/MD_ElementFlux.cpp:54-80
// (avgY_sf: src/Equations/Equations.cpp:8-51).  isf / nsf already clamped at 0.
__device__ __forceinline__ double edge_surface(double isf, double zs, double nsf, double zs_n, double depression,
                                               double dist, double B, double rough) {
    // written without branches (selects only) so that the three edges of a cell, which are independent, are
    // interleaved by the scheduler instead of being three serial sqrt / cbrt / multiply chains
    const double h1 = zs + isf, h2 = zs_n + nsf;
    const double dh = (isf + zs) - (nsf + zs_n);
    double ym = (h1 > h2) ? ((isf > depression) ? isf : 0.) : ((nsf > depression) ? nsf : 0.);
    ym = dmin(ym, kMAXYSURF);
    const double s = SHUD_DIVS(dh, dist);
    const bool off = (ym <= 0.) || (s > 0 && isf <= 0) || (s < 0 && nsf <= 0);
    const double yms = off ? 1.0 : ym;                    // keep the dead lanes on harmless arguments
    const double as = (s > 0) ? s : -s;
    const double c = cbrt(yms);
    const double q = SHUD_DIVS(fsqrt(off ? 1.0 : as) * (yms * B) * (c * c), rough);   // ManningEquation, |S| form
    return off ? 0. : ((s > 0) ? q : -1.0 * q);
}

// groundwater flux through one edge, fun_Ele_sub, src/ModelData/MD_ElementFlux.cpp:107-138 (before fu_Sub).
// (y_n, z_n) = neighbour head pair: (uYgw, z_bottom) of a cell, or (yLakeStg, bathymetry.yi[0]) of a lake.
__device__ __forceinline__ double edge_sub(double ygw, double zb, double y_n, double z_n, double kh, double kh_n,
                                           double dist, double B) {
    const double dh = (ygw + zb) - (y_n + z_n);
    const bool off = (dh > 0. && ygw <= 0.02) || (dh < 0. && y_n <= 0.02);
    const double ym = (dmax(ygw, 0.) + dmax(y_n, 0.)) * .5;  // avgY_gw, Equations.cpp:52-56
    const double grad = SHUD_DIVS(dh, dist);
    const double km = 0.5 * (kh + kh_n);
    const double q = km * grad * ym * B;
    return off ? 0. : q;
}

}  // namespace shud
