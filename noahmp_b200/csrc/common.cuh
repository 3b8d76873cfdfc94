// noahmp_b200 -- device-side constants, table layout and small helpers.
//
// Constants are the fp32 literals of core/module_noahmp_const.f90:14-35; table slots are the C-ABI
// packing of include/noahmp_b200.h.  Column arrays keep the reference's layer numbering through the
// index macros below: snow+soil arrays (-2:4) -> [k+2], snow arrays (-2:0) -> [k+2], soil (1:4) -> [k-1].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmp {

constexpr int NSOIL = 4;   // core/module_noahmp_global.f90:11
constexpr int NSNOW = 3;   // core/module_noahmp_global.f90:13

constexpr float MPE    = 1.0E-6f;
constexpr float GRAV   = 9.80616f;
constexpr float SB     = 5.67E-8f;
constexpr float RGAS   = 8.3144598f;
constexpr float KARMAN = 0.40f;
constexpr float TFRZ   = 273.15f;
constexpr float HSUB   = 2.8440E6f;
constexpr float HVAP   = 2.5104E6f;
constexpr float HFUS   = 0.3336E6f;
constexpr float CWAT   = 4.188E6f;
constexpr float CICE   = 2.094E6f;
constexpr float CPAIR  = 1004.64f;
constexpr float TKWAT  = 0.6f;
constexpr float TKICE  = 2.2f;
constexpr float RAIR   = 287.04f;
constexpr float RVAP   = 461.269f;
constexpr float DENWAT = 1000.0f;
constexpr float DENICE = 917.0f;

#define IX7(k) ((k) + 2)   // Fortran (-2:4)
#define IX3(k) ((k) + 2)   // Fortran (-2:0)
#define IX4(k) ((k) - 1)   // Fortran (1:4)

// ---- plane indices (include/noahmp_b200.h) ------------------------------------------------------
enum {
  S_ALBOLD = 0, S_SNEQVO, S_EAH, S_FWET, S_CANLIQ, S_CANICE, S_TV, S_TG, S_QSNOW, S_ISNOW, S_SNOWH, S_SNEQV,
  S_ZWT, S_WA, S_WT, S_TAUSS, S_STC = 16, S_SOILWAT = 23, S_SMC = 27, S_ZSNSO = 31, S_SNICE = 38, S_SNLIQ = 41,
  NSTATE_CORE = 44,
  S_QSFC = 44, S_TAH, S_CM, S_CH, S_WSLAKE, S_LFMASS, S_RTMASS, S_STMASS, S_WOOD, S_STBLCP, S_FASTCP, S_LAI,
  S_SAI, NSTATE = 57
};
enum { F_SFCTMP = 0, F_SFCPRS, F_UU, F_VV, F_Q2, F_SOLDN, F_LWDN, F_PRCP, F_COSZ, NFORCING = 9 };
enum { C_LAT = 0, C_SHDFAC, C_TBOT, C_SHDMAX, NSTATIC_F = 4 };
enum {
  D_FSA = 0, D_FSR, D_FIRA, D_FSH, D_SSOIL, D_FCEV, D_FGEV, D_FCTR, D_ECAN, D_ETRAN, D_EDIR, D_TRAD, D_TGB, D_TGV,
  D_T2MV, D_T2MB, D_Q2V, D_Q2B, D_RUNSRF, D_RUNSUB, D_APAR, D_PSN, D_SAV, D_SAG, D_FSNO, D_NEE, D_GPP, D_NPP,
  D_FVEG, D_ALBEDO, D_QSNBOT, D_PONDING, D_PONDING1, D_PONDING2, D_RSSUN, D_RSSHA, D_BGAP, D_WGAP, D_CHV, D_CHB,
  D_EMISSI, D_SHG, D_SHC, D_SHB, D_EVG, D_EVB, D_GHV, D_GHB, D_IRG, D_IRC, D_IRB, D_TR, D_EVC, D_CHLEAF, D_CHUC,
  D_CHV2, D_CHB2, D_FPICE, D_ERRWAT, D_ERRSW, D_ERRENG, D_ITER, D_FLAGS,
  // accumulated over every step since the caller last zeroed them (mm): the planes are read, added to and written back
  D_ACCPRCP = 63, D_ACCECAN, D_ACCETRAN, D_ACCEDIR, D_SFCRNOFF, D_UGDRNOFF, NDIAG = 69
};
enum {
  B_NCELL = 0, B_NFLAGGED, B_ERRWAT_SUM, B_ERRWAT_ABS, B_ERRWAT_MAXABS, B_FSA, B_FSR, B_FIRA, B_FSH, B_LH, B_SSOIL,
  B_PRCP_DT, B_ET_DT, B_RUNOFF_DT, B_SNEQV, B_SOILWATER, NBUDGET = 16
};
enum {
  FLAG_ERRSW = 1u << 0, FLAG_ERRENG = 1u << 1, FLAG_FIRE = 1u << 2, FLAG_HCAN = 1u << 3, FLAG_ZLVL = 1u << 4,
  FLAG_OPTVEG = 1u << 5, FLAG_NAN = 1u << 6, FLAG_BADTYPE = 1u << 7
};

// ---- packed parameter tables ----------------------------------------------------------------------
enum {
  G_CSOIL = 0, G_ZBOT, G_CZIL, G_DKREF, G_KDTREF, G_FRZK, G_TIMEAN, G_FSATMAX, G_MLTFCT, G_Z0SNO, G_SSI, G_SWEMAX,
  G_ALBICE = 12, G_ALBLAKE = 14, G_OMEGAS = 16, G_BETADS = 18, G_BETAIS, G_EMSSOIL, G_EMSLAKE, G_NSLOPE = 22,
  G_SLOPE = 23, NGEN = 53
};
enum { P_BEXP = 0, P_SMCMAX, P_SMCREF, P_SMCWLT, P_PSISAT, P_DKSAT, P_DWSAT, P_QUARTZ, P_KDT, P_FRZX, NSOILP = 10 };
enum { K_ALBSAT = 0 /*vis,nir*/, K_ALBDRY = 2, NCOLORP = 4 };
enum {
  V_XL = 0, V_RHOL = 1, V_RHOS = 3, V_TAUL = 5, V_TAUS = 7, V_CANWMXP = 9, V_DLEAF, V_Z0MVT, V_HVT, V_HVB, V_DEN,
  V_RCROWN, V_CWPVT, V_LAI12M = 17, V_SAI12M = 29, V_SLA = 41, V_DILEFC, V_DILEFW, V_FRAGR, V_LTOVRC, V_WRRAT,
  V_WDPOOL, V_TDLEF, V_RGL = 49, V_HS, V_KC25, V_AKC, V_KO25, V_AKO, V_VCMX25, V_AVCMX, V_BP, V_RSMAX, V_RSMIN, V_MP,
  V_QE25, V_AQE, V_RMF25, V_RMS25, V_RMR25, V_FOLNMX, V_TOPT, V_TMIN, V_ARM, V_MRP, NVEGF = 71
};
enum { VI_NROOT = 0, VI_C3C4, NVEGI = 2 };

constexpr int MAXSOILTYP = 30, MAXCOLOR = 20, MAXVEGTYP = 27;

// One contiguous block in global memory (ctx-owned), staged into shared memory by every CTA.
struct Tables {
  float gen[NGEN];
  float soil[MAXSOILTYP * NSOILP];
  float color[MAXCOLOR * NCOLORP];
  float veg[MAXVEGTYP * NVEGF];
  int   vegi[MAXVEGTYP * NVEGI];
  int   flags[5];   // ISURBAN ISWATER ISBARREN ISICE ISEGBLF
  int   nsoiltyp, ncolor, nvegtyp;
};
static_assert(sizeof(Tables) % 4 == 0, "Tables is copied word by word");

// ---- option policies --------------------------------------------------------------------------------
// The reference keeps 12 module-level integers (core/module_noahmp_global.f90:17-74).  Here they are a
// POD passed to the kernel; the shipped option set (run/case.nml:29-37 + defaults) is also available as
// a compile-time policy so every option branch folds away.
struct OptRuntime {
  int veg, crs, btr, run, sfc, frz, inf, rad, alb, snf, tbot, stc;
};
struct OptDefault {
  static constexpr int veg = 1, crs = 1, btr = 1, run = 1, sfc = 1, frz = 1, inf = 1, rad = 1, alb = 2, snf = 1,
                       tbot = 1, stc = 1;
  __host__ __device__ OptDefault() {}
  __host__ __device__ OptDefault(const OptRuntime&) {}
};

template <class Opt> struct OptTraits;
template <> struct OptTraits<OptDefault> { static constexpr bool runtime = false; };
template <> struct OptTraits<OptRuntime> { static constexpr bool runtime = true; };

// ---- per-launch scalars ---------------------------------------------------------------------------
struct RunConst {
  float dt;
  float zsoil[4];
  float zref, co2_frac, o2_frac, foln;
  int   yearlen;
  double julian0;
};

__device__ __forceinline__ float fmin2(float a, float b) { return a < b ? a : b; }   // Fortran MIN
__device__ __forceinline__ float fmax2(float a, float b) { return a > b ? a : b; }   // Fortran MAX
__device__ __forceinline__ float sq(float x) { return x * x; }
__device__ __forceinline__ float cube(float x) { return x * x * x; }                 // x**3
__device__ __forceinline__ float pow4(float x) { float y = x * x; return y * y; }    // x**4
__device__ __forceinline__ float pow5(float x) { float y = x * x; return y * y * x; }// x**5

// ---- libm layer -----------------------------------------------------------------------------------
// Default: CUDA's fp32 libm (expf/logf/powf/...: max error 1-4 ulp).  With -DNMP_CR_LIBM every call is
// evaluated in fp64 and rounded once to fp32, i.e. correctly rounded except with probability ~2^-29.  The
// CPU oracle has the same switch (oracle/src/common.h, ORACLE_CR_LIBM); in that mode the two independent
// implementations must agree bit for bit, which is how the arithmetic of this kernel is verified
// (tests/test_parity_gpu.py::test_strict_bit_exact).
#ifdef NMP_CR_LIBM
#define NMP_LIBM1(name, dname) __device__ __forceinline__ float m_##name(float x) { return (float)dname((double)x); }
__device__ __forceinline__ float m_powf(float x, float y) { return (float)pow((double)x, (double)y); }
#else
#define NMP_LIBM1(name, dname) __device__ __forceinline__ float m_##name(float x) { return name(x); }
#ifdef NMP_ACCURATE_POW
// under --use_fast_math `powf` becomes exp2(y*log2(x)) on the MUFU units (relative error ~|y log2 x| 2^-22);
// libdevice's full-precision routine is still reachable by its internal name
extern "C" __device__ float __nv_powf(float, float);
__device__ __forceinline__ float m_powf(float x, float y) { return __nv_powf(x, y); }
#else
__device__ __forceinline__ float m_powf(float x, float y) { return powf(x, y); }
#endif
#endif
// Correctly rounded quotient whatever -prec-div says.  The soil-moisture planes go volumetric -> mass -> volumetric
// every step (func:4360/4423, 6527/6636); with the approximate reciprocal of the fast build x*d*rcp(d) is x(1+e) with the
// SAME e each step (d = 1000 mm for layer 4 gives e = 4.75e-8), which compounds to 4e-4 of the deep soil moisture in a
// year of hourly steps.  Those few conversions therefore divide exactly in every build.
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// the soil hydraulic functions (func:6386-6455) raise SMC/SMCMAX to powers up to 2b+3 ~ 26 three to six times
// per step; their error accumulates in the soil water, so they can be kept accurate on their own
#if defined(NMP_ACCURATE_SOILPOW) && !defined(NMP_CR_LIBM)
extern "C" __device__ float __nv_powf(float, float);
__device__ __forceinline__ float m_soilpowf(float x, float y) { return __nv_powf(x, y); }
#else
__device__ __forceinline__ float m_soilpowf(float x, float y) { return m_powf(x, y); }
#endif
#if defined(NMP_FAST_ALGEBRA) && !defined(NMP_CR_LIBM)
extern "C" __device__ float __nv_powf(float, float);
__device__ __forceinline__ float m_accpowf(float x, float y) { return __nv_powf(x, y); }
#endif
NMP_LIBM1(expf, exp)
NMP_LIBM1(logf, log)
NMP_LIBM1(log10f, log10)
NMP_LIBM1(tanhf, tanh)
NMP_LIBM1(atanf, atan)
NMP_LIBM1(tanf, tan)
NMP_LIBM1(acosf, acos)
NMP_LIBM1(cosf, cos)
#undef NMP_LIBM1

// Phase barriers: the step is ~16k SASS instructions of mostly straight-line code, far beyond the
// instruction caches, and ncu shows the warps starving on instruction fetch (stall_no_inst).  A CTA-wide
// barrier at every phase boundary keeps all warps of the CTA inside the same few-KB code window, so a line
// fetched for one warp is reused by the others.  Only placed where control flow is convergent.
// Each site has a number (source order: 0-2 column_step, 3-14 energy, 15-18 water) so that a placement experiment
// can drop single barriers with -DNMP_SYNC_MASK=0x... (bit n = site n on); the shipped builds keep them all.
#ifndef NMP_SYNC_MASK
#define NMP_SYNC_MASK 0xffffffffu
#endif
#ifdef NMP_PHASE_BARRIERS
#define NMP_PHASE_SYNC_AT(n) do { if ((NMP_SYNC_MASK >> (n)) & 1u) __syncthreads(); } while (0)
#else
#define NMP_PHASE_SYNC_AT(n) ((void)0)
#endif
// Loops whose trip count depends on the data (canopy Newton loop, stomata bisection, soil sub-steps) can be
// run CTA-uniform (-DNMP_UNIFORM_LOOPS: every thread executes the loop control, one barrier per pass).
// Measured on B200 this LOSES ~15%: the whole CTA then waits for its slowest lane (some lane of 512-768
// almost always needs all 20 passes), so the default keeps these loops warp-divergent and barrier-free.
#if defined(NMP_PHASE_BARRIERS) && defined(NMP_UNIFORM_LOOPS)
#define NMP_ANY(pred) (__syncthreads_or((pred) ? 1 : 0) != 0)
#define NMP_LOOP_SYNC() __syncthreads()
#else
#define NMP_ANY(pred) (pred)
#define NMP_LOOP_SYNC() ((void)0)
#endif
#ifndef NMP_BLOCK
#define NMP_BLOCK 128
#endif

// Compile-time loops.  Layer loops must be unrolled in the FRONT END: with `#pragma unroll` the optimiser
// first propagates `k == ISNOW+1` into the body as a dynamic index and the column arrays fall out of
// registers into local memory.  Here k is a constant expression inside the body.
template <int I> struct IntC { static constexpr int value = I; };
template <int A, int B, class F>
__device__ __forceinline__ void static_for(F&& f) {   // A, A+1, ..., B
  if constexpr (A <= B) {
    f(IntC<A>{});
    static_for<A + 1, B>(f);
  }
}
template <int A, int B, class F>
__device__ __forceinline__ void static_for_rev(F&& f) {   // A, A-1, ..., B
  if constexpr (A >= B) {
    f(IntC<A>{});
    static_for_rev<A - 1, B>(f);
  }
}

// ---- bounds-checked debug build (-DNMP_BOUNDS; compute-sanitizer is not available on the GPU pool) ------------------
// Every data-dependent index of the step - pick(), the snow-pack scratch of combine / divide, the month indices of
// phenology, the table rows - goes through NMP_CHK.  In the debug build an index outside [lo, hi] is counted, its site is
// recorded (bit `site` of g_bounds_sites; read back with noahmp_debug_bounds) and the access is clamped instead of being
// made; in every other build the macro is the index itself.  Sites: 0 pick() of a layer array, 1 combine's scratch,
// 2 divide's scratch, 3 month index of phenology, 4 table row of a type index, 5 root-layer count, 6 water-table layer.
#ifdef NMP_BOUNDS
__device__ unsigned long long g_bounds_hits;
__device__ unsigned int g_bounds_sites;
__device__ __forceinline__ int bounds_chk(int idx, int lo, int hi, int site) {
  if (idx < lo || idx > hi) {
    atomicAdd(&g_bounds_hits, 1ull);
    atomicOr(&g_bounds_sites, 1u << site);
    return idx < lo ? lo : hi;
  }
  return idx;
}
#define NMP_CHK(idx, lo, hi, site) nmp::bounds_chk((idx), (lo), (hi), (site))
#else
#define NMP_CHK(idx, lo, hi, site) (idx)
#endif

// select element `idx` of a small register array without dynamic indexing
template <int N>
__device__ __forceinline__ float pick(const float (&a)[N], int idx, int site = 0) {
  idx = NMP_CHK(idx, 0, N - 1, site);
  float r = a[0];
  static_for<1, N - 1>([&](auto i_) { constexpr int i = decltype(i_)::value; r = (idx == i) ? a[i] : r; });
  return r;
}

// -ZSOIL(NROOT), the depth of the root zone (func:332-335, 1119-1140, 6700).  The reference evaluates it only inside loops
// over layers 1..NROOT, so a type without roots (NROOT = 0: water, ice) never reads ZSOIL(0); here the value is formed
// ahead of the (unrolled) loop and is unused in that case.  The debug build checks NROOT against 0..NSOIL.
__device__ __forceinline__ float root_depth(const RunConst& rc, int NROOT) {
  NROOT = NMP_CHK(NROOT, 0, NSOIL, 5);
  return -pick(rc.zsoil, NROOT > 0 ? NROOT - 1 : 0);
}

}  // namespace nmp
