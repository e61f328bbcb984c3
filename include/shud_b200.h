/* shud_b200.h - C ABI of the B200-native SHUD hot path.
 *
 * Drop-in boundary for the reference's CVODE right-hand side and the N_Vector
 * arithmetic CVODE/SPGMR runs on the same state vectors.
 * Plain pointers and sizes only; the library behind it is hand-written sm_100a
 * CUDA (shud_up_b200/csrc).  There is NO CPU fallback: every entry point returns
 * SHUD_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * State vector layout (reference src/Model/Macros.hpp:21-25, blocked):
 *     y = [ Ysurf[Ne] | Yunsat[Ne] | Ygw[Ne] | Yriv[Nr] | Ylake[Nl] ],  NY = 3 Ne + Nr + Nl
 * Units: metres, minutes (reference src/ModelData/MD_readin.cpp:290-349).
 * Indices in nabr/lakenabr/iLake/down/iEle/iRiv are 1-BASED exactly as the reference
 * holds them (0 = none, negative = special); [3][Ne] arrays are edge-major:
 * a[j*Ne + i] is edge j of cell i.
 */
#ifndef SHUD_B200_H
#define SHUD_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SHUD_OK 0
#define SHUD_ERR_NO_DEVICE (-1) /* no CUDA device / extension unusable: never falls back to the CPU */
#define SHUD_ERR_ARG (-2)
#define SHUD_ERR_CUDA (-3)
/* device-side checks mirror the reference's myexit() codes (src/Model/Macros.hpp:77-82) */
#define SHUD_ERRNAN 10    /* CheckNANi / CheckNonNegative, src/Equations/functions.cpp:90-103,148-154 */
#define SHUD_ERRDATAIN 13 /* effKH out of range, src/Equations/Equations.cpp:129-132 */
#define SHUD_ERRRIVBC 1   /* unknown river 'down' code, src/ModelData/MD_RiverFlux.cpp:55-57 (exit(1)) */
#define SHUD_ERR_P2P_TIMEOUT 99 /* device error word: a neighbour's halo flag did not arrive within 5 s (peer-to-peer exchange) */

/* ---- static model data: what Model_Data::initialize() leaves behind, flattened
 * AoS -> SoA once after initialize()+LoadIC() (reference src/Model/shud.cpp:50-67).
 * Field names are the reference's (src/classes/Element.hpp, River.hpp, Lake.hpp). ---- */
typedef struct shud_mesh {
    int32_t Ne, Nr, Ns, Nl;
    int32_t close_boundary; /* Control_Data::CloseBoundary, src/classes/Model_Control.hpp:164 */
    int32_t lakeon;         /* global lakeon, src/Model/shud.cpp:30 */
    /* per cell [Ne] */
    const double *area, *z_surf, *z_bottom, *depression, *AquiferDepth, *Sy;
    const double *infD, *infKsatV, *macKsatV, *hAreaF, *ThetaS, *ThetaR, *ThetaFC, *Beta;
    const double *KsatH, *KsatV, *macKsatH, *macD, *geo_vAreaF;
    const double *VegFrac, *ImpAF, *WetlandLevel, *RootReachLevel, *Rough, *QSS;
    /* per cell and edge [3][Ne] */
    const double *edge, *Dist2Nabor, *Dist2Edge, *avgRough;
    const int32_t *nabr, *lakenabr;
    /* per cell [Ne] */
    const int32_t *iLake, *iBC, *iSS;
    /* cell centroids (Triangle::x,y, src/classes/Element.hpp:32-33): used ONLY to order cells along a
     * Hilbert curve for memory locality; may be NULL (then the given order is kept) */
    const double *x, *y;
    /* per reach [Nr] */
    const double *riv_Length, *riv_BedSlope, *riv_depth, *riv_BottomWidth, *riv_bankslope;
    const double *riv_avgRough, *riv_Dist2DownStream, *riv_KsatH, *riv_BedThick, *riv_zbank;
    const int32_t *riv_down, *riv_BC, *riv_toLake;
    /* per river-element segment [Ns] */
    const int32_t *seg_iEle, *seg_iRiv;
    const double *seg_length, *seg_Cwr;
    /* per lake [Nl]; bathymetry table of lake l is rows lake_bathy_ptr[l] .. lake_bathy_ptr[l+1]-1 */
    const double *lake_zmin;
    const int32_t *lake_NumEleLake;
    const int32_t *lake_bathy_ptr; /* [Nl+1] */
    const double *lake_bathy_yi, *lake_bathy_ai;
} shud_mesh;

/* ---- values that change only between CVode() calls (after updateforcing()+ET(),
 * reference src/Model/shud.cpp:106-109): one upload per forcing step. ---- */
typedef struct shud_forcing {
    const double *qEleNetPrep, *qPotEvap, *qPotTran, *t_lai, *fu_Surf, *fu_Sub, *qElePrep; /* [Ne] */
    const double *qEleE_IC; /* [Ne] read AND rewritten by the RHS (src/ModelData/MD_ET.cpp:370,381) */
    /* boundary-condition values already looked up (tsd_*BC.getX is a step function of the
     * ring pointer, src/classes/TimeSeriesData.cpp:270-273).  NULL = no such BC anywhere. */
    const double *ele_yBC, *ele_QBC; /* [Ne] */
    const double *riv_yBC, *riv_qBC; /* [Nr] */
} shud_forcing;

/* ---- flux arrays the reference's Print_Ctrl reads (src/ModelData/MD_initialize.cpp:258-342);
 * filled by the *_diag RHS variant.  Any pointer may be NULL (skipped). ---- */
typedef struct shud_diag {
    double *qEleInfil, *qEleExfil, *qEleRecharge;            /* [Ne] */
    double *qEs, *qEu, *qEg, *qTu, *qTg;                     /* [Ne] */
    double *qEleTrans, *qEleEvapo, *qEleETA, *iBeta;         /* [Ne] */
    double *u_effKH, *u_satn;                                /* [Ne] */
    double *QeleSurf, *QeleSub;                              /* [3][Ne] */
    double *QeleSurfTot, *QeleSubTot, *Qe2r_Surf, *Qe2r_Sub; /* [Ne] */
    double *QsegSurf, *QsegSub;                              /* [Ns] */
    double *QrivSurf, *QrivSub, *QrivUp, *QrivDown;          /* [Nr] */
    double *y2LakeArea, *QLakeSurf, *QLakeSub, *QLakeRivIn, *QLakeRivOut, *qLakeEvap, *qLakePrcp; /* [Nl] */
} shud_diag;

/* ---- one partition of a larger mesh (multi-GPU).  The shud_mesh of a partition
 * holds its OWNED cells, reaches, segments and lakes; a neighbour index nabr in Ne+1 .. Ne+Nhalo names halo
 * cell (nabr-Ne-1): a cell owned by another partition whose state arrives by halo exchange before each RHS.
 * Only what an edge flux needs from the far side is kept for a halo cell.  Restrictions of this version:
 * halo cells are plain land cells (no head BC, not lake), and reaches / lakes are not cut by the partition. ---- */
typedef struct shud_halo {
    int32_t Nhalo;
    const double *z_surf, *z_bottom;                               /* [Nhalo] geometry of the far cell */
    const double *AquiferDepth, *macD, *macKsatH, *geo_vAreaF, *KsatH; /* [Nhalo] its effKH parameters */
    /* Cut river trees: the LAST n_ghost_cells cells and the LAST n_ghost_reaches reaches of the shud_mesh are GHOSTS -
     * bank cells of this partition's reaches that another partition owns (no edges; their vertical role and segment
     * fluxes are evaluated here from the owner's (Ysurf, Yunsat, Ygw)), and reaches on this partition's banks / up- or
     * downstream of its reaches / flowing into its lakes that another partition owns (stage from the owner).  Their
     * states arrive with the halo exchange (shud_b200_exchange_plan_items), their entries of y are never read, their
     * entries of ydot are 0.  Both 0: whole river trees per partition, as before. */
    int32_t n_ghost_cells, n_ghost_reaches;
} shud_halo;

typedef struct shud_ctx shud_ctx;

/* Build the device-resident SoA mirror (cells reordered for locality, CSR gathers for
 * segment->cell, segment->reach, reach->downstream, bank-edge->lake) and the CUDA graph.
 * `device` is the CUDA ordinal.  Replaces: nothing in the reference - it is the one-time
 * export after Model_Data::initialize() (src/Model/shud.cpp:51). */
int shud_b200_create(const shud_mesh *mesh, int device, shud_ctx **out);
/* Same, for one partition with halo cells (halo may be NULL = shud_b200_create). */
int shud_b200_create_partition(const shud_mesh *mesh, const shud_halo *halo, int device, shud_ctx **out);
/* Device buffer [Nhalo][2] = (Ysurf, Ygw) of every halo cell: the receive buffer of the halo exchange,
 * registered once and read by the kernels of every following RHS. */
int shud_b200_set_halo_state_dev(shud_ctx *ctx, const double *halo_state_dev);
/* Pack (Ysurf, Ygw) of `n` owned cells (device-order ids in idx_dev) into out_dev[2k], out_dev[2k+1]:
 * the send side of the halo exchange (same pair layout, so a peer's message lands in place). */
int shud_b200_pack_halo_dev(shud_ctx *ctx, const double *y_dev, const int32_t *idx_dev, int32_t n, double *out_dev);
void shud_b200_destroy(shud_ctx *ctx);
int64_t shud_b200_ny(const shud_ctx *ctx);
/* the CUDA stream (cudaStream_t) every call of this context is ordered on */
void *shud_b200_stream(shud_ctx *ctx);

/* Upload one forcing step (host pointers).  Replaces the implicit hand-over of
 * qEleNetPrep.. after Model_Data::updateforcing()+ET() (src/Model/shud.cpp:106-109). */
int shud_b200_set_forcing(shud_ctx *ctx, const shud_forcing *f);
/* Carried state: Ele[i].u_satn left by the previous call.
 * shud_b200_prime computes it from y the way Model_Data::updateforcing does through
 * _Element::updateElement (src/ModelData/MD_ET.cpp:14-19, src/classes/Element.cpp:347-373). */
int shud_b200_prime(shud_ctx *ctx, const double *y_host);
int shud_b200_set_carried(shud_ctx *ctx, const double *u_satn_host);
int shud_b200_get_carried(shud_ctx *ctx, double *u_satn_host, double *qEleE_IC_host);

/* Device vectors handed to the *_dev entry points and to the N_Vector ops are in DEVICE ORDER:
 * same blocked layout, cells and reaches renumbered for locality.  These two kernels convert a
 * device-resident vector between the reference order and the device order (both pointers on the
 * device, src != dst); shud_b200_perm() exposes the maps (device id -> 0-based reference id). */
int shud_b200_to_device_order(shud_ctx *ctx, const double *ref_order_dev, double *dev_order_dev);
int shud_b200_from_device_order(shud_ctx *ctx, const double *dev_order_dev, double *ref_order_dev);
/* host vector in the reference's blocked order <-> device vector in device order (one copy + one permutation kernel;
 * both synchronise): the host mirror of the device N_Vector (SetIC2Y writes, src/ModelData/MD_initialize.cpp:117-135) */
int shud_b200_upload_ref(shud_ctx *ctx, const double *y_host_ref, double *y_dev);
int shud_b200_download_ref(shud_ctx *ctx, const double *y_dev, double *y_host_ref);
/* Model_Data::summary(N_Vector) (src/ModelData/MD_update.cpp:190-216), the read-back the driver does every SolverStep:
 * the device vector (device order) lands in y_host_ref [NY] in the reference's blocked order, with the groundwater
 * head of iBC > 0 cells and the stage of BC > 0 reaches replaced by their boundary values.  Synchronises. */
int shud_b200_summary_dev(shud_ctx *ctx, const double *y_dev, double *y_host_ref);
int shud_b200_perm(const shud_ctx *ctx, int32_t *cell_perm /*[Ne]*/, int32_t *reach_perm /*[Nr]*/);

/* The RHS.  Replaces int f(double t, N_Vector y, N_Vector ydot, void *MD)
 * (src/Model/f.hpp:12, src/Model/f.cpp:2-32) = f_update + f_loop + f_applyDY
 * (src/ModelData/MD_update.cpp:102-189, MD_f.cpp:9-50, MD_f.cpp:52-215).
 * _dev: y/ydot are device pointers, asynchronous on shud_b200_stream().
 * plain: y/ydot are host pointers; copies in, runs, copies out, synchronises and
 * returns the device error word (0, or SHUD_ERRNAN / SHUD_ERRDATAIN / SHUD_ERRRIVBC). */
int shud_b200_rhs_dev(shud_ctx *ctx, double t, const double *y_dev, double *ydot_dev);
int shud_b200_rhs(shud_ctx *ctx, double t, const double *y_host, double *ydot_host);
/* The difference-quotient evaluation of CVLS' Jv (cvLsDQJtimes, inside SUNLinSol_SPGMR), the perturbation folded into
 * the pre-pass of the RHS: ytemp = sigma (v ./ ewt) + y0 - the arithmetic of shud_nv_dq_perturb - and
 * ydot = f(t, ytemp), without a vector kernel of its own and without reading ytemp back for effKH.  Single domain
 * (a partition's pre-pass carries the halo exchange: SHUD_ERR_ARG there; perturb, then shud_b200_rhs_dev).
 * ss_dev != NULL: v is an unnormalised Krylov vector with squared 2-norm *ss_dev (device memory, e.g. the result of
 * the Gram-Schmidt sweep's last reduction); the direction is (1 / sqrt(ss)) v, stored to vnorm_dev (!= v_dev) when
 * that is not NULL - the solver's normalisation pass done on the way. */
int shud_b200_dq_foldable(const shud_ctx *ctx); /* 1: shud_b200_rhs_dq_dev serves this context (single domain) */
int shud_b200_rhs_dq_dev(shud_ctx *ctx, double t, double sigma, const double *v_dev, const double *ewt_dev,
                         const double *y0_dev, double *ytemp_dev, double *ydot_dev, const double *ss_dev,
                         double *vnorm_dev);
/* The RHS of a partition in two parts, so that the halo exchange overlaps the bulk of the work: _interior needs no
 * exchanged data (effKH of the owned cells + every tile of cells that sees no halo cell); _boundary (after the
 * exchange has landed in the halo state buffer) does the halo effKH, the remaining tiles and the river/lake kernel.
 * interior followed by boundary == shud_b200_rhs_dev.  `halo_stream` (cudaStream_t) is the stream on which the
 * exchange completes: the halo-dependent tiles are enqueued there (behind an event for the owned cells' effKH) and
 * run beside the interior tiles; the context stream then waits for them before the river/lake kernel.  NULL = the
 * context stream (the exchange was ordered on it; everything serial). */
int shud_b200_rhs_interior_dev(shud_ctx *ctx, double t, const double *y_dev, double *ydot_dev);
/* The halo exchange driven by the library itself (replaces the MPI/NCCL calls a host code would place around the
 * RHS, SURVEY.md section 8(e) "grouped ncclSend/ncclRecv per neighbour pair"): NCCL is opened at run time from
 * `nccl_lib` (path of libnccl.so.2; NULL = the loader's default) - the library does not link it.
 *   shud_b200_comm_unique_id  rank 0 obtains the 128-byte ncclUniqueId; the host distributes it (MPI_Bcast, ...)
 *   shud_b200_comm_init       every rank, same id; creates the communicator and the exchange stream
 *   shud_b200_exchange_plan   neighbours of this partition: rank, number of owned cells sent, number of halo cells
 *                             received (halo cells are numbered by (owner rank, global id), so the receives land
 *                             contiguously in halo order; sum recv_count == Nhalo); `send_cells`: 0-based local
 *                             reference ids of the cells sent, concatenated by peer, each peer's in the order the
 *                             peer numbers them as its halo cells.  Allocates the send / halo buffers.
 *   shud_b200_rhs_exchange_dev  one f(): pack -> sends/receives on the exchange stream -> interior part beside
 *                             them -> boundary part.  Asynchronous on the context stream like shud_b200_rhs_dev. */
int shud_b200_comm_unique_id(const char *nccl_lib, void *id128);
int shud_b200_comm_init(shud_ctx *ctx, const char *nccl_lib, const void *id128, int rank, int world);
int shud_b200_exchange_plan(shud_ctx *ctx, int npeers, const int32_t *peer_rank, const int32_t *send_count,
                            const int32_t *recv_count, const int32_t *send_cells);
int shud_b200_rhs_exchange_dev(shud_ctx *ctx, double t, const double *y_dev, double *ydot_dev);
/* The general exchange plan (peer-to-peer path; required when the partition has ghosts): per neighbour three groups of
 * doubles travel - kind 0 halo-cell pairs (Ysurf, Ygw), kind 1 ghost-cell triples (Ysurf, Yunsat, Ygw), kind 2
 * ghost-reach stages.  send_count / recv_count: [npeers][3] doubles per (neighbour, kind); the receives of a kind land
 * in neighbour order, which must be the order the halo cells / ghost cells / ghost reaches are numbered in (by owner,
 * then global id).  send_items: flat 0-based indices into THIS partition's blocked vector in reference-local order
 * ([Ysurf | Yunsat | Ygw | Yriv | Ylake] over its local cells and reaches), grouped by (neighbour, kind). */
int shud_b200_exchange_plan_items(shud_ctx *ctx, int npeers, const int32_t *peer_rank, const int32_t *send_count,
                                  const int32_t *recv_count, const int32_t *send_items);
/* Peer-to-peer halo exchange (one process per GPU of one node, NVLink / NVSwitch): instead of NCCL sends / receives the
 * pack kernel stores every boundary cell's state straight into the halo buffer of the partition that needs it (the
 * neighbours' buffers are mapped through CUDA IPC) and releases one flag per neighbour; the receiver's boundary tiles
 * start behind a flag wait.  The step is then a CUDA graph of plain kernels.  After shud_b200_exchange_plan:
 *   shud_b200_p2p_export   fills a SHUD_P2P_BLOB_BYTES blob describing this rank's halo buffer (IPC handle, who lands where)
 *   (the host gathers the blobs of all ranks, rank order: MPI_Allgather / torch.distributed.all_gather)
 *   shud_b200_p2p_connect  maps the neighbours' buffers; a host barrier must follow before the first exchange.
 * shud_b200_rhs_exchange_dev then uses this path (SHUD_P2P=0 in the environment keeps the NCCL path).  Contexts living
 * in one process (tests) are connected by pointer, no IPC. */
#define SHUD_P2P_BLOB_BYTES 512
int shud_b200_p2p_export(shud_ctx *ctx, int rank, void *blob);
int shud_b200_p2p_connect(shud_ctx *ctx, int rank, int world, const void *blobs);
/* Scalar allreduce over the same communicator for the distributed N_Vector reductions:
 * `vals` are n HOST doubles reduced in place over all ranks; op 0 sum, 1 max, 2 min.  Synchronises the context stream. */
int shud_b200_allreduce(shud_ctx *ctx, double *vals, int n, int op);
/* The same on doubles already in DEVICE memory, enqueued on `stream` without copy or synchronisation (ctx = the shud_ctx):
 * the allreduce hook of a distributed N_Vector workspace (shud_nv_ws_set_allreduce, include/shud_nvector.h). */
int shud_b200_allreduce_dev(void *ctx, double *dev_vals, int n, int op, void *stream);
/* After shud_b200_p2p_connect: the mailboxes of all ranks for the in-kernel allreduce of the device vector's reductions
 * (boxes[0..*nranks), for shud_nv_ws_set_peer_allreduce; *nranks = 0 when a rank's block could not be mapped or the
 * ranks share a process).  N_VSetDistributed_ShudB200 on the library's communicator installs them by itself. */
int shud_b200_p2p_mailboxes(shud_ctx *ctx, int *nranks, int *rank, void **boxes);
/* ---- land-surface step on the device ----
 * Replaces the per-cell loops of Model_Data::updateforcing / tReadForcing (src/ModelData/MD_ET.cpp:14-281:
 * lapse-rate temperature, terrain-radiation factor, Penman-Monteith potential evaporation / transpiration) and
 * Model_Data::ET (MD_ET.cpp:282-342: snow and interception buckets, net precipitation), which the reference runs
 * on the host once per ET step (src/Model/shud.cpp:106-109).  The results are written straight into the device
 * arrays the RHS reads (what shud_b200_set_forcing would upload: qEleNetPrep, qPotEvap, qPotTran, t_lai, fu_Surf,
 * fu_Sub, qEleE_IC, and the lake means of qPotEvap / qElePrep), so the 9*Ne doubles no longer cross PCIe.
 * What stays on the host is O(stations + classes) per step: the time-series lookups and solarPosition(). */
typedef struct shud_land {
    int32_t nforc, nlc, nmf;                         /* forcing stations, land-cover classes, melt-factor classes */
    const int32_t *iForc, *iLC, *iMF;                /* [Ne] 1-based ids, Element.hpp:49-52 */
    const double *Albedo, *FixPressure, *windH;      /* [Ne] */
    const double *nx, *ny, *nz;                      /* [Ne] unit surface normal (terrain radiation), Element.hpp:38 */
    const double *forc_z;                            /* [nforc] station elevation; -9999 = none (no lapse-rate shift) */
    double cPrep, cTemp, cLAItsd, cMF, cETP, cISmax; /* calibration multipliers, ModelConfigure.hpp (globalCal) */
    int32_t radiation_is_net;                        /* RADIATION_INPUT_MODE == SWNET (MD_ET.cpp:208-214) */
    int32_t terrain_radiation;                       /* TERRAIN_RADIATION */
    int32_t cryosphere;                              /* CRYOSPHERE: frozen-soil factors fu_Surf / fu_Sub (MD_ET.cpp:301-311) */
    double rad_factor_cap, rad_cosz_min;
    /* CRYOSPHERE = 1: window lengths [days] and thresholds of the two running means of the daily mean air
     * temperature (calib_frozen, ModelConfigure.hpp:43-52; _AccTemp, AccTemperature.hpp) */
    double FT_surf_day, FT_surf_max, FT_surf_min, FT_sub_day, FT_sub_max, FT_sub_min;
} shud_land;

typedef struct shud_land_step {
    const double *forc;      /* [nforc][5] prcp, temp, rh, wind, rn of each station's current interval (i_prcp..i_rn) */
    const double *lai, *mf;  /* [nlc], [nmf] tsd_LAI / tsd_MF value of each class at t */
    int32_t tsr_n;           /* solar samples of the forcing interval (host: solarPosition(), MD_ET.cpp:74-138);
                                < 0: the interval start is not finite (factor 0, MD_ET.cpp:66-67) */
    const double *tsr_sx, *tsr_sy, *tsr_sz, *tsr_wdt;  /* [tsr_n] */
    double tsr_den;
    double dt_min;           /* tnext - t of ET(t, tnext) */
    double t;                /* t of ET(t, tnext) [min]: clock of the frozen-soil accumulators (CRYOSPHERE = 1) */
} shud_land_step;

typedef struct shud_land_out { /* host arrays [Ne] in reference order, any may be NULL */
    double *qElePrep, *qPotEvap, *qPotTran, *qEleETP, *t_lai, *t_temp, *t_mf, *qEleNetPrep, *qEleE_IC, *fu_Surf,
        *fu_Sub, *rn_factor, *yEleSnow, *yEleIS;
} shud_land_out;

int shud_b200_land_create(shud_ctx *ctx, const shud_land *L);
int shud_b200_land_set_state(shud_ctx *ctx, const double *yEleSnow, const double *yEleIS); /* host [Ne] */
int shud_b200_land_step(shud_ctx *ctx, const shud_land_step *S); /* asynchronous on the context stream */
int shud_b200_land_get(shud_ctx *ctx, const shud_land_out *out);  /* synchronises */
/* ---- ingest and checkpoint, host side ----
 * shud_b200_mesh_save / _load: the shud_mesh SoA as ONE binary file (header + 64-byte aligned arrays) instead of the
 * reference's ~10 text files parsed with strtold per field (src/classes/TabularData.cpp:27-55,
 * src/ModelData/MD_readin.cpp:192-236).  _load reads the payload with a single read into one block (returned in
 * *block, release with shud_b200_mesh_free after shud_b200_create) and points *out into it.
 * shud_b200_format_ic: Model_Data::PrintInit (src/ModelData/MD_update.cpp:268-299), byte-identical text
 * ("<prj>.cfg.ic.update"); y in the reference's blocked order, yEleIS / yEleSnow may be NULL (zeros).
 * shud_b200_write_ic: the same from a DEVICE vector in device order; the canopy / snow buckets come from the
 * device land-surface step when it is in use, else zeros; like PrintInit after summary(), head-BC cells and
 * stage-BC reaches print their boundary value (shud_b200_summary_dev). */
int shud_b200_mesh_save(const char *path, const shud_mesh *m);
int shud_b200_mesh_load(const char *path, shud_mesh *out, void **block);
void shud_b200_mesh_free(void *block);
int shud_b200_format_ic(const char *path, double t, int32_t Ne, int32_t Nr, int32_t Nl, const double *yEleIS,
                        const double *yEleSnow, const double *y);
int shud_b200_write_ic(shud_ctx *ctx, const char *path, double t, const double *y_dev);
/* the reverse of shud_b200_format_ic (restart): y [3 Ne + Nr + Nl] blocked, reference order; t, yEleIS, yEleSnow may
 * be NULL.  SHUD_ERR_ARG when the file does not match the sizes. */
int shud_b200_read_ic(const char *path, int32_t Ne, int32_t Nr, int32_t Nl, double *t, double *yEleIS, double *yEleSnow,
                      double *y);
/* number of 128-cell tiles in each part (interior + boundary = ceil(Ne/128)) */
int shud_b200_tile_counts(const shud_ctx *ctx, int *n_interior, int *n_boundary);
int shud_b200_rhs_boundary_dev(shud_ctx *ctx, double t, const double *y_dev, double *ydot_dev, void *halo_stream);
/* One launch of the RHS sequence alone (stage 0 effKH pre-pass, 1 cell kernel, 2 river+lake kernel):
 * for per-kernel CUDA-event timing and ncu; shud_b200_rhs_dev == stages 0,1,2 in order. */
int shud_b200_rhs_stage_dev(shud_ctx *ctx, int stage, const double *y_dev, double *ydot_dev);
/* Same arithmetic, and additionally stores every flux array of shud_diag on the device. */
int shud_b200_rhs_diag_dev(shud_ctx *ctx, double t, const double *y_dev, double *ydot_dev);
/* Download the flux arrays left by the last shud_b200_rhs_diag_dev (host pointers). */
int shud_b200_get_diag(shud_ctx *ctx, const shud_diag *out);
/* Device-side output accumulation.  Replaces the host loop of
 * Print_Ctrl::PrintData (src/classes/Model_Control.cpp:930-962): every SolverStep `buffer[i] += *PrintVar[i]`,
 * and at the end of the output interval `buffer[i] *= tau / NumUpdate`, write, reset.  _accumulate adds the flux
 * arrays left by the last shud_b200_rhs_diag_dev into device-resident buffers (NumUpdate++), no PCIe traffic;
 * _flush downloads buffer * (tau / NumUpdate) into the host arrays of `out` (reference order) and resets. */
int shud_b200_output_accumulate(shud_ctx *ctx);
int shud_b200_output_flush(shud_ctx *ctx, double tau, const shud_diag *out, int32_t *num_update);
/* Synchronise and read the device error word: code in the return value, offending
 * 1-based cell/reach id in *where (may be NULL).  Mirrors myexit(code)
 * (src/Equations/functions.cpp:10-36) without killing the process. */
int shud_b200_check(shud_ctx *ctx, int32_t *where);
/* number of kernels one shud_b200_rhs_dev launches (for bench.py's gpu_launches) */
int shud_b200_launches_per_rhs(const shud_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* SHUD_B200_H */
