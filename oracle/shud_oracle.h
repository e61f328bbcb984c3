/* TEST INFRASTRUCTURE ONLY - see shud_oracle.c.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this. */
#ifndef SHUD_ORACLE_H
#define SHUD_ORACLE_H
#include "shud_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* One reference f() call (src/Model/f.cpp:2-32) on SoA inputs.
 *   u_satn   [Ne] in/out : Ele[i].u_satn left by the previous call
 *   qEleE_IC [Ne] in/out : rewritten by f_etFlux (src/ModelData/MD_ET.cpp:370,381)
 *   diag     may be NULL, and any pointer inside may be NULL
 *   nthreads 1 = the reference's serial order exactly; >1 = same arithmetic with
 *            `omp parallel for` on the cell / segment / reach loops (CPU timing baseline)
 * returns 0, or the reference's exit code (10 NaN / negative ET, 13 effKH range, 1 river BC). */
int shud_oracle_rhs(const shud_mesh *m, const shud_forcing *f, double *u_satn, double *qEleE_IC,
                    const double *y, double *ydot, const shud_diag *diag, int nthreads);
/* Ele[i].u_satn as Model_Data::updateforcing -> _Element::updateElement leaves it
 * (src/ModelData/MD_ET.cpp:14-19, src/classes/Element.cpp:347-373). */
void shud_oracle_prime(const shud_mesh *m, const double *y, double *u_satn);
/* One land-surface step: Model_Data::updateforcing -> tReadForcing (src/ModelData/MD_ET.cpp:14-281) followed by
 * Model_Data::ET (MD_ET.cpp:282-342), per cell, on the SoA inputs of shud_land / shud_land_step.
 *   z_surf, VegFrac, iLake : the shud_mesh members;  yEleSnow, yEleIS [Ne] in/out;  out: any pointer may be NULL
 * returns 0, or 10 where the reference exits (CheckNonZero of the aerodynamic resistance, NaN qPotTran). */
int shud_oracle_land_step(const shud_mesh *m, const shud_land *L, const shud_land_step *S, double *yEleSnow,
                          double *yEleIS, const shud_land_out *out, double *cryo);
/* `cryo` (CRYOSPHERE = 1, else NULL): the two _AccTemp accumulators of every cell (AccTemperature.hpp), carried
 * between calls.  8 scalars [Time_start, N_of_day, size_surf, head_surf, size_sub, head_sub, -, -] then per-cell
 * arrays T_AccDay[Ne], ACC_surf[Ne], ACC_sub[Ne], ring_surf[FT_surf_day][Ne], ring_sub[FT_sub_day][Ne].
 * Initial state: all zeros except Time_start = -9999 (shud_oracle_cryo_size doubles). */
long shud_oracle_cryo_size(const shud_mesh *m, const shud_land *L);
/* how many scratch doubles per call the oracle allocates (informational) */
const char *shud_oracle_version(void);
#ifdef __cplusplus
}
#endif
#endif
