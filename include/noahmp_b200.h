/* noahmp_b200 -- C ABI of the B200-native Noah-MP column time step.
 *
 * This library fills the slot the reference leaves empty: core/module_noahmp_engine.f90:5-10 declares
 * `noahmp_init` and `noahmp_run` with no arguments and no body, and nothing in the reference ever calls
 * the one-column physics routine `noahmp_sflx` (core/module_noahmp_func.f90:66-91).  Here `noahmp_run`
 * advances ALL columns of a structure-of-arrays batch on one GPU; per column the semantics are those of
 * one `noahmp_sflx` call per step.  Plain C types only; all pointers named *_dev are CUDA device
 * pointers owned by the caller (e.g. torch tensors' data_ptr()), all others are host pointers.
 *
 * Return convention (replaces the reference's `stop` in utils:43-53 / wrf_error_fatal sites):
 *   0 = OK, < 0 = API misuse or CUDA error (text via noahmp_last_error), > 0 = number of columns whose
 *   error-flag word is non-zero (only reported by calls that compute it).  Nothing aborts or throws.
 */
#ifndef NOAHMP_B200_H
#define NOAHMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct noahmp_ctx noahmp_ctx;

/* ---- plane counts (see noahmp_layout) -------------------------------------------------------- */
#define NOAHMP_NSTATE      57  /* 44 core planes + 13 extended (SURVEY App. A groups S and s)      */
#define NOAHMP_NSTATE_CORE 44
#define NOAHMP_NFORCING     9  /* SFCTMP SFCPRS UU VV Q2 SOLDN LWDN PRCP COSZ  (func:72-74)        */
#define NOAHMP_NSTATIC_F    4  /* LAT SHDFAC TBOT SHDMAX                                           */
#define NOAHMP_NDIAG       69  /* 58 outputs of func:82-91 + ERRWAT ERRSW ERRENG ITER FLAGS + 6 accumulations */
#define NOAHMP_NBUDGET     16
#define NOAHMP_NGEN        53
#define NOAHMP_NSOILP      10
#define NOAHMP_NCOLORP      4
#define NOAHMP_NVEGF       71
#define NOAHMP_NVEGI        2

/* error-flag bits: one per fatal site of the reference */
#define NOAHMP_FLAG_ERRSW   (1u << 0) /* func:688-710  */
#define NOAHMP_FLAG_ERRENG  (1u << 1) /* func:712-721  */
#define NOAHMP_FLAG_FIRE    (1u << 2) /* func:1286-1292 */
#define NOAHMP_FLAG_HCAN    (1u << 3) /* func:2726-2738 */
#define NOAHMP_FLAG_ZLVL    (1u << 4) /* func:3412-3415 */
#define NOAHMP_FLAG_OPTVEG  (1u << 5) /* func:375-378  */
#define NOAHMP_FLAG_NAN     (1u << 6) /* non-finite prognostic state (added) */
#define NOAHMP_FLAG_BADTYPE (1u << 7) /* lutyp/sltyp/slptyp/ISC outside the tables (added): column left untouched */

/* per-run scalars: the arguments of noahmp_sflx that do not vary per column (func:67-74) */
typedef struct noahmp_run_params {
  double julian0;   /* JULIAN at the first step; step s uses (float)(julian0 + s*dt/86400)        */
  float  dt;        /* DT [s]                                                                      */
  float  zsoil[4];  /* ZSOIL(1:4) [m], negative, layer bottoms                                     */
  float  zref;      /* reference height of the forcing above the canopy/displacement height [m]   */
  float  co2_frac;  /* CO2AIR = co2_frac * SFCPRS                                                  */
  float  o2_frac;   /* O2AIR  = o2_frac  * SFCPRS                                                  */
  float  foln;      /* FOLN [%]                                                                    */
  int    yearlen;   /* YEARLEN [days]                                                              */
} noahmp_run_params;

/* ---- lifecycle: replaces engine:5-6 (noahmp_init) ---------------------------------------------- */
int noahmp_init(noahmp_ctx** ctx, int device);
int noahmp_finalize(noahmp_ctx* ctx);
/* Run on this CUDA stream (a cudaStream_t passed as void*); NULL = the legacy default stream. */
int noahmp_set_stream(noahmp_ctx* ctx, void* cuda_stream);
const char* noahmp_last_error(const noahmp_ctx* ctx);   /* ctx == NULL: text of the last failed noahmp_init */

/* ---- options: replaces noahmp_set_options, core/module_noahmp_global.f90:77-112, same order:
 * veg, crs, btr, run, sfc, frz, inf, rad, alb, snf, tbot, stc.  Out-of-range values are rejected
 * (< 0) except opt_veg, which the reference itself checks per column (func:375-378 -> FLAG_OPTVEG). */
int noahmp_set_options(noahmp_ctx* ctx, const int opts[12]);

/* ---- parameter tables: replaces noahmp_{gen,soil,veg}_param_readptable (gen_param:51, soil_param:31,
 * veg_param:77).  The host parses the .TBL files and hands packed arrays:
 *   gen    [NGEN]             slots of enum order CSOIL ZBOT CZIL DKREF KDTREF FRZK TIMEAN FSATMAX MLTFCT
 *                             Z0SNO SSI SWEMAX ALBICE(2) ALBLAKE(2) OMEGAS(2) BETADS BETAIS EMSSOIL
 *                             EMSLAKE NSLOPE SLOPE(30)
 *   soil   [nsoiltyp][NSOILP] BEXP SMCMAX SMCREF SMCWLT PSISAT DKSAT DWSAT QUARTZ KDT FRZX
 *   color  [ncolor][NCOLORP]  ALBSAT_VIS ALBSAT_NIR ALBDRY_VIS ALBDRY_NIR
 *   veg_f  [nvegtyp][NVEGF]   RAD(9) PHYS(8) LAI12M(12) SAI12M(12) DVEG(8) PHOTO(22)
 *   veg_i  [nvegtyp][NVEGI]   NROOT C3C4
 *   flags  [5]                ISURBAN ISWATER ISBARREN ISICE ISEGBLF */
int noahmp_set_tables(noahmp_ctx* ctx, const float* gen, int ngen, const float* soil, int nsoiltyp,
                      const float* color, int ncolor, const float* veg_f, const int* veg_i, int nvegtyp,
                      const int flags[5]);

/* ---- the hot path: replaces engine:8-10 (noahmp_run) + the missing caller loop of func:66 --------
 * Advance ncol columns by nsteps time steps on the context's stream (asynchronous).
 *   state_dev    [NSTATE][ncol]  fp32 planes, in/out (ISNOW as int32 bits)
 *   forcing_dev  step s at forcing_dev + s*forcing_step_stride, each [NFORCING][ncol]
 *   static_f_dev [NSTATIC_F][ncol]; static_i_dev [ncol] packed types:
 *                lutyp | sltyp<<8 | slptyp<<16 | ISC<<24 | (IST-1)<<28 | ICE<<29
 *   diag_dev     [NDIAG][ncol] or NULL: planes 0..62 hold the values of the LAST step; planes 63..68 (ACCPRCP ACCECAN
 *                ACCETRAN ACCEDIR SFCRNOFF UGDRNOFF, mm) are accumulations - every step's precipitation, canopy
 *                evaporation, transpiration, ground evaporation, surface and subsurface runoff x dt is ADDED to what
 *                the plane holds, so they run across launches until the caller zeroes them (output cadence,
 *                offline/noahmp_config.py:83-84 output_frequency)
 *   budget_dev   [NBUDGET] fp64 sums accumulated (+=) over all columns and steps, or NULL          */
int noahmp_run(noahmp_ctx* ctx, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
               const float* forcing_dev, int64_t forcing_step_stride, const float* static_f_dev,
               const int32_t* static_i_dev, float* diag_dev, double* budget_dev);

/* Same call with HOST buffers: copies inputs to the device, runs, copies state (and diag/budget if
 * non-NULL) back and synchronises.  forcing is [nsteps][NFORCING][ncol].  Returns the number of
 * flagged columns (>= 0) or < 0.  This is the reference-facing entry a host driver binds. */
int64_t noahmp_run_host(noahmp_ctx* ctx, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state,
                        const float* forcing, const float* static_f, const int32_t* static_i, float* diag,
                        double* budget);

/* Streaming variant for a time-marching driver: the column state and the static planes STAY on the device
 * (uploaded once, like a model's weights), only the forcing arrives from the host.  Forcing of launch k+1 is
 * copied H2D on a private copy stream into a double buffer while launch k computes (`steps_per_launch` model
 * steps per launch, >= 1).  forcing_host is [nsteps][NFORCING][ncol] and should be page-locked for the copies
 * to overlap.  budget_host (nullable) receives the 16 budget sums of this call.  Synchronises before returning
 * the number of column-launches that raised a flag (>= 0; a column flagged in m launches counts m times) or
 * < 0 on error. */
int64_t noahmp_advance_host(noahmp_ctx* ctx, const noahmp_run_params* rp, int64_t ncol, int nsteps,
                            float* state_dev, const float* forcing_host, const float* static_f_dev,
                            const int32_t* static_i_dev, double* budget_host, int steps_per_launch);

/* ---- records-driven launches: forcing ingest fused into the step kernel -------------------
 * The namelist gives forcing records at `input_frequency` and a model step of `interval_seconds`
 * (offline/noahmp_config.py:81-82,98).  Instead of preparing nine forcing planes per step in a pass of its own
 * (noahmp_prepare_forcing below) the step kernel can read the RECORDS: per step it interpolates the two bracketing
 * records and computes COSZ from the column's position - the same round-to-nearest arithmetic, bit for bit - so a
 * record crosses PCIe / HBM once however many model steps it serves, and 8 planes travel instead of 9. */
#define NOAHMP_MAX_RECORD_STEPS 32   /* model steps of one records-driven kernel launch */
typedef struct noahmp_step_forcing {
  int   rec0, rec1;        /* the two bracketing records: forcing = rec0 + weight * (rec1 - rec0), plane by plane     */
  float weight;            /* (t - t0) / (t1 - t0) in [0,1]; 0: rec1 is not read                                      */
  float sd, cd, ch0, sh0;  /* sin / cos of the solar declination, cos / sin of the hour angle at longitude 0          */
} noahmp_step_forcing;

/* noahmp_run with the forcing given as records: record r at records_dev + r*record_stride, each [8][ncol] (SFCTMP SFCPRS UU
 * VV Q2 SOLDN LWDN PRCP); steps[nsteps] is a HOST array; geo_dev [4][ncol] = sin(lat) cos(lat) sin(lon) cos(lon).
 * nsteps beyond NOAHMP_MAX_RECORD_STEPS is cut into several launches.  Otherwise as noahmp_run. */
int noahmp_run_records(noahmp_ctx* ctx, const noahmp_run_params* rp, int64_t ncol, int nsteps, float* state_dev,
                       const float* records_dev, int64_t record_stride, const noahmp_step_forcing* steps,
                       const float* geo_dev, const float* static_f_dev, const int32_t* static_i_dev, float* diag_dev,
                       double* budget_dev);

/* Streaming variant (the time-marching driver's call): state, static and geo planes stay on the device, the records
 * arrive from the host.  records_host is [nrec][8][ncol] (page-locked for the copies to overlap); steps[nsteps] index
 * into it.  Every record is copied H2D ONCE, on a private copy stream into a ring of device slots, while earlier
 * launches compute (`steps_per_launch` model steps per launch, 1..NOAHMP_MAX_RECORD_STEPS).  The ring holds up to 24
 * records (a day of hourly ones; never less than 4 x steps_per_launch, never more than a quarter of the free device
 * memory), so that the copies can run ahead of the kernel through the expensive hours of the day.  Returns like
 * noahmp_advance_host. */
int64_t noahmp_advance_records_host(noahmp_ctx* ctx, const noahmp_run_params* rp, int64_t ncol, int nsteps,
                                    float* state_dev, const float* records_host, int nrec,
                                    const noahmp_step_forcing* steps, const float* geo_dev, const float* static_f_dev,
                                    const int32_t* static_i_dev, double* budget_host, int steps_per_launch);

/* Zero suppression of record planes on their way through noahmp_advance_records_host.  Short-wave radiation is zero
 * wherever it is night and precipitation wherever it is dry: before a record crosses the host link, host threads pack
 * the candidate planes (bit q of plane_mask = plane q of a record; default SOLDN | PRCP = 0xA0, 0 = off) into a 32-column
 * occupancy word + running count per group and the non-zero words themselves, and the step kernel's forcing load reads
 * that form directly.  Lossless and bit-identical: only the bit pattern 0 is suppressed, a plane that is more than 80 %
 * non-zero travels as it is.  host_threads <= 0: as many as the calling thread may run on, at most 8.  The reference has
 * no counterpart (it has no forcing ingest at all; offline/noahmp_config.py:81-82 only names the record cadence). */
int noahmp_set_record_packing(noahmp_ctx* ctx, unsigned plane_mask, int host_threads);

/* The packer by itself (host only, no device needed; exported for tests): plane[ncol] -> occupancy[ceil(ncol/32)],
 * before[ceil(ncol/32)] (non-zero words in earlier groups), values[<= ncol].  Returns the number of non-zero words. */
int64_t noahmp_pack_plane(const float* plane, int64_t ncol, uint32_t* occupancy, uint32_t* before, float* values,
                          int host_threads);

/* Bytes the noahmp_advance_* entry points of this context have copied host -> device so far. */
int64_t noahmp_h2d_bytes(const noahmp_ctx* ctx);
/* Wall time (seconds) noahmp_advance_records_host has spent packing record planes on the host so far. */
double noahmp_pack_seconds(const noahmp_ctx* ctx);

/* ---- synthetic forcing generator (benchmark support; no counterpart in the reference) ------------
 * Fills forcing_dev [NFORCING][ncol] for one step with the hash-based generator of
 * noahmp_b200/synth.py (bit-identical).  sincos_dev [2][ncol] = sin(lat), cos(lat).
 * coef[17] = synth.synth_params_array(); step_scalars[4] = sd, cdch, t_diur, t_seas. */
int noahmp_synth_forcing(noahmp_ctx* ctx, int64_t ncol, int64_t col0, uint32_t seed, int step,
                         const float* coef17, const float* step_scalars4, const float* sincos_dev,
                         float* forcing_dev);

/* the same for re-ordered columns: ids_dev [ncol] int64 = each column's id in the generator's numbering */
int noahmp_synth_forcing_ids(noahmp_ctx* ctx, int64_t ncol, const int64_t* ids_dev, uint32_t seed, int step,
                             const float* coef17, const float* step_scalars4, const float* sincos_dev,
                             float* forcing_dev);

/* ---- forcing preparation (the caller side of the path; the reference has the namelist
 * fields `input_frequency` / `interval_seconds` but no code): time-interpolate two forcing records to the model
 * step and compute COSZ.
 *   rec0_dev, rec1_dev [8][ncol]  SFCTMP SFCPRS UU VV Q2 SOLDN LWDN PRCP at the two bracketing record times
 *   weight                        (t - t0) / (t1 - t0) in [0,1]
 *   geo_dev [4][ncol]             sin(lat) cos(lat) sin(lon) cos(lon)
 *   scal4                         sin(decl), cos(decl), cos(H0), sin(H0)   (H0 = hour angle at longitude 0)
 *   forcing_dev [NFORCING][ncol]  out: the 9 planes noahmp_run reads */
int noahmp_prepare_forcing(noahmp_ctx* ctx, int64_t ncol, const float* rec0_dev, const float* rec1_dev, float weight,
                           const float* geo_dev, const float scal4[4], float* forcing_dev);

/* ---- introspection ------------------------------------------------------------------------------ */
int noahmp_layout(int what); /* 0 NSTATE 1 NSTATE_CORE 2 NFORCING 3 NSTATIC_F 4 NDIAG 5 NGEN 6 NSOILP
                                7 NCOLORP 8 NVEGF 9 NVEGI 10 sizeof(noahmp_run_params) 11 NBUDGET */
const char* noahmp_state_plane_name(int i);
const char* noahmp_forcing_plane_name(int i);
const char* noahmp_diag_plane_name(int i);
const char* noahmp_budget_name(int i);
/* kernels launched by this context since noahmp_init (for the benchmark's gpu_launches claim) */
int64_t noahmp_launch_count(const noahmp_ctx* ctx);
/* registers / local-memory bytes / static smem of the step kernel variant that `noahmp_run` would use */
int noahmp_kernel_info(noahmp_ctx* ctx, int info[4]);
/* Debug instrumentation (no counterpart in the reference, which is compiled without bounds checks): in a library built
 * with -DNMP_BOUNDS every data-dependent index of the step kernel is range-checked on the device (csrc/common.cuh,
 * NMP_CHK); violations are counted and clamped, never made.  Returns 1 and the counts since the last reset (bit n of
 * `sites` = site n was hit) in such a build, 0 (counts zero) in every other build, < 0 on error. */
int noahmp_debug_bounds(noahmp_ctx* ctx, unsigned long long* hits, unsigned int* sites, int reset);
/* launches one thread that selects element `idx` of a 4-element array through the checked helper (proves the checker
 * counts); 1 in a -DNMP_BOUNDS build, 0 otherwise */
int noahmp_debug_bounds_selftest(noahmp_ctx* ctx, int idx);

#ifdef __cplusplus
}
#endif
#endif /* NOAHMP_B200_H */
