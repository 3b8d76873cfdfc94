// TEST INFRASTRUCTURE ONLY -- CPU parity oracle for the Noah-MP column time step.
//
// PARITY PIN: the reference (wk1984/noahmp-1) ships no golden vectors and no tests, and no Fortran compiler exists
// in this image, so the Fortran cannot be run as it is.  Instead oracle/f90_to_cpp.py translates the reference's own
// source (core/module_noahmp_*.f90, read where it lies) statement by statement into C++ under oracle/_ref/ (not in
// git), and tests/golden/ref_*.npz hold that translation's outputs for every workload and every value of every option
// (tools/make_reference_golden.py).  This hand-written restatement reproduces all of them BIT FOR BIT
// (tests/test_reference_golden.py).  What the pin does not cover: a real Fortran compiler's code generation (it may
// legally differ, e.g. GCC with -fno-math-errno rewrites cos(atan(x)), see Makefile) and the .TBL readers.
// It is further checked by the reference's own conservation asserts (func:688-731) and leaf-routine known answers.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build or load this.
// The product (noahmp_b200/) never includes, links or calls anything under oracle/.
//
// Conventions: every routine keeps the Fortran name, argument meaning and operation order of
// the reference routine it cites; Fortran index ranges are kept through the A7/A3/A4 accessors.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

// ---- libm layer: with -DORACLE_CR_LIBM every fp32 libm call is evaluated in fp64 and rounded once
// (correctly rounded except with probability ~2^-29).  The CUDA library has the same switch
// (NMP_CR_LIBM); in that mode the two independent implementations must agree bit for bit.  The default
// build keeps glibc's fp32 functions, i.e. what a gfortran build of the reference would call.
#ifdef ORACLE_CR_LIBM
namespace noahmp_oracle_cr {
static inline float f1(double (*fn)(double), float x) { return (float)fn((double)x); }
static inline float powf_(float x, float y) { return (float)std::pow((double)x, (double)y); }
}  // namespace noahmp_oracle_cr
#define expf(x) noahmp_oracle_cr::f1(std::exp, (x))
#define logf(x) noahmp_oracle_cr::f1(std::log, (x))
#define log10f(x) noahmp_oracle_cr::f1(std::log10, (x))
#define tanhf(x) noahmp_oracle_cr::f1(std::tanh, (x))
#define atanf(x) noahmp_oracle_cr::f1(std::atan, (x))
#define tanf(x) noahmp_oracle_cr::f1(std::tan, (x))
#define acosf(x) noahmp_oracle_cr::f1(std::acos, (x))
#define cosf(x) noahmp_oracle_cr::f1(std::cos, (x))
#define powf(x, y) noahmp_oracle_cr::powf_((x), (y))
#endif

namespace noahmp_oracle {

// ---- dimensions: core/module_noahmp_global.f90:9-13
constexpr int NBAND = 2;
constexpr int NSOIL = 4;
constexpr int NSNOW = 3;

// ---- physical constants: core/module_noahmp_const.f90:14-35 (fp32 literals)
constexpr float MPE    = 1.0E-6f;
constexpr float GRAV   = 9.80616f;
constexpr float SB     = 5.67E-8f;
constexpr float RGAS   = 8.3144598f;
constexpr float KARMAN = 0.40f;
constexpr float TFRZ   = 273.15f;
constexpr float TTRI   = 273.16f;
constexpr float HSUB   = 2.8440E6f;
constexpr float HVAP   = 2.5104E6f;
constexpr float HFUS   = 0.3336E6f;
constexpr float CWAT   = 4.188E6f;
constexpr float CICE   = 2.094E6f;
constexpr float CPAIR  = 1004.64f;
constexpr float TKWAT  = 0.6f;
constexpr float TKICE  = 2.2f;
constexpr float TKAIR  = 0.023f;
constexpr float RAIR   = 287.04f;
constexpr float RVAP   = 461.269f;
constexpr float DENWAT = 1000.0f;
constexpr float DENICE = 917.0f;

// ---- Fortran-indexed small arrays
struct A7 {  // (-NSNOW+1:NSOIL) = (-2:4)
  float v[7];
  float& operator()(int i) { return v[i + 2]; }
  const float& operator()(int i) const { return v[i + 2]; }
};
struct A3 {  // (-NSNOW+1:0) = (-2:0)
  float v[3];
  float& operator()(int i) { return v[i + 2]; }
  const float& operator()(int i) const { return v[i + 2]; }
};
struct A4 {  // (1:NSOIL)
  float v[4];
  float& operator()(int i) { return v[i - 1]; }
  const float& operator()(int i) const { return v[i - 1]; }
};
struct A2 {  // (1:2) radiation bands
  float v[2];
  float& operator()(int i) { return v[i - 1]; }
  const float& operator()(int i) const { return v[i - 1]; }
};
struct I7 {  // integer (-2:4)
  int v[7];
  int& operator()(int i) { return v[i + 2]; }
  const int& operator()(int i) const { return v[i + 2]; }
};

// ---- options: core/module_noahmp_global.f90:17-74, set order global:77-112
struct Options {
  int opt_veg, opt_crs, opt_btr, opt_run, opt_sfc, opt_frz, opt_inf, opt_rad, opt_alb, opt_snf,
      opt_tbot, opt_stc;
};

// ---- parameter tables (1-based type index kept; slot 0 unused)
constexpr int MSLOPETYP = 30;  // gen_param:8
constexpr int MSLTYP = 30;     // soil_param:7
constexpr int MSLCOL = 20;     // soil_param:8
constexpr int MLUTYP = 27;     // veg_param:8

struct Tables {
  // gen_param:12-47
  float LK_SLOPE[MSLOPETYP + 1];
  float KK_CSOIL, KK_ZBOT, KK_CZIL, KK_DKREF, KK_KDTREF, KK_FRZK, KK_TIMEAN, KK_FSATMAX, KK_MLTFCT,
      KK_Z0SNO, KK_SSI, KK_SWEMAX;
  float KK_ALBICE[3], KK_ALBLAKE[3], KK_OMEGAS[3];
  float KK_BETADS, KK_BETAIS, KK_EMSSOIL, KK_EMSLAKE;
  // soil_param:13-28
  int nsltyp, nsoilcol;
  float LK_BEXP[MSLTYP + 1], LK_SMCMAX[MSLTYP + 1], LK_SMCREF[MSLTYP + 1], LK_SMCWLT[MSLTYP + 1],
      LK_PSISAT[MSLTYP + 1], LK_DKSAT[MSLTYP + 1], LK_DWSAT[MSLTYP + 1], LK_QUARTZ[MSLTYP + 1],
      LK_KDT[MSLTYP + 1], LK_FRZX[MSLTYP + 1];
  float LK_ALBSAT[3][MSLCOL + 1], LK_ALBDRY[3][MSLCOL + 1];  // [band][color]
  // veg_param:13-74
  int nlutyp;
  int ISURBAN, ISWATER, ISBARREN, ISICE, ISEGBLF;
  float LK_XL[MLUTYP + 1];
  float LK_RHOL[3][MLUTYP + 1], LK_RHOS[3][MLUTYP + 1], LK_TAUL[3][MLUTYP + 1], LK_TAUS[3][MLUTYP + 1];
  int LK_NROOT[MLUTYP + 1];
  float LK_CANWMXP[MLUTYP + 1], LK_DLEAF[MLUTYP + 1], LK_Z0MVT[MLUTYP + 1], LK_HVT[MLUTYP + 1],
      LK_HVB[MLUTYP + 1], LK_DEN[MLUTYP + 1], LK_RCROWN[MLUTYP + 1], LK_CWPVT[MLUTYP + 1];
  float LK_SAI12M[13][MLUTYP + 1], LK_LAI12M[13][MLUTYP + 1];  // [month][type]
  float LK_SLA[MLUTYP + 1], LK_DILEFC[MLUTYP + 1], LK_DILEFW[MLUTYP + 1], LK_FRAGR[MLUTYP + 1],
      LK_LTOVRC[MLUTYP + 1], LK_WRRAT[MLUTYP + 1], LK_WDPOOL[MLUTYP + 1], LK_TDLEF[MLUTYP + 1];
  float LK_RGL[MLUTYP + 1], LK_HS[MLUTYP + 1], LK_RSMAX[MLUTYP + 1], LK_RSMIN[MLUTYP + 1],
      LK_TOPT[MLUTYP + 1];
  int LK_C3C4[MLUTYP + 1];
  float LK_KC25[MLUTYP + 1], LK_AKC[MLUTYP + 1], LK_KO25[MLUTYP + 1], LK_AKO[MLUTYP + 1],
      LK_VCMX25[MLUTYP + 1], LK_AVCMX[MLUTYP + 1], LK_BP[MLUTYP + 1], LK_MP[MLUTYP + 1],
      LK_QE25[MLUTYP + 1], LK_AQE[MLUTYP + 1], LK_FOLNMX[MLUTYP + 1], LK_TMIN[MLUTYP + 1],
      LK_RMF25[MLUTYP + 1], LK_RMS25[MLUTYP + 1], LK_RMR25[MLUTYP + 1], LK_ARM[MLUTYP + 1],
      LK_MRP[MLUTYP + 1];
};

// ---- error flag bits: one per fatal site of the reference; never abort
enum : uint32_t {
  FLAG_ERRSW   = 1u << 0,  // func:688-710  |SWDOWN-(FSA+FSR)| > 0.01
  FLAG_ERRENG  = 1u << 1,  // func:712-721  surface energy residual > 0.01
  FLAG_FIRE    = 1u << 2,  // func:1286-1292 emitted longwave <= 0
  FLAG_HCAN    = 1u << 3,  // func:2726-2738 HCAN <= ZPD
  FLAG_ZLVL    = 1u << 4,  // func:3412-3415 ZLVL <= ZPD
  FLAG_OPTVEG  = 1u << 5,  // func:375-378  unknown opt_veg
  FLAG_NAN     = 1u << 6,  // non-finite prognostic state after the step (added; not in reference)
};

// Fortran integer power by repeated multiplication, as gfortran expands x**n for small n.
static inline float pow2i(float x) { return x * x; }
static inline float pow3i(float x) { return x * x * x; }
static inline float pow4i(float x) { float x2 = x * x; return x2 * x2; }
static inline float pow5i(float x) { float x2 = x * x; return x2 * x2 * x; }

static inline float fmin2(float a, float b) { return a < b ? a : b; }  // Fortran MIN
static inline float fmax2(float a, float b) { return a > b ? a : b; }  // Fortran MAX

// ---- column "call frame": everything noahmp_sflx takes (func:66-91), per column
struct ColumnIO {
  // --- global/config (G)
  int   YEARLEN;
  float JULIAN, DT;
  A4    ZSOIL;
  float CO2AIR, O2AIR, FOLN;
  // --- per-column static (C)
  float LAT, SHDFAC, SHDMAX, TBOT, ZLVL;
  int   slptyp, sltyp, lutyp, ICE, IST, ISC;
  // --- forcing (F)
  float COSZ, SFCTMP, SFCPRS, PSFC, UU, VV, Q2, SOLDN, LWDN, PRCP;
  A3    FICEOLD;
  // --- state (S + s)
  float ALBOLD, SNEQVO;
  A7    STC;
  A4    soilwat, SMC;
  float TAH, EAH, FWET, CANLIQ, CANICE, TV, TG, QSFC, QSNOW;
  int   ISNOW;
  A7    ZSNSO;
  float snowh, sneqv;
  A3    SNICE, SNLIQ;
  float ZWT, WA, WT, WSLAKE, LFMASS, RTMASS, STMASS, WOOD, STBLCP, FASTCP, LAI, SAI, CM, CH, TAUSS;
  // --- diagnostics (D), order of func:82-91
  float FSA, FSR, FIRA, FSH, SSOIL, FCEV, FGEV, FCTR, ECAN, ETRAN, EDIR, TRAD, TGB, TGV, T2MV, T2MB,
      Q2V, Q2B, runsrf, runsub, APAR, PSN, SAV, SAG, FSNO, NEE, GPP, NPP, fveg, ALBEDO, QSNBOT,
      PONDING, PONDING1, PONDING2, RSSUN, RSSHA, BGAP, WGAP, CHV, CHB, EMISSI, SHG, SHC, SHB, EVG,
      EVB, GHV, GHB, IRG, IRC, IRB, TR, EVC, CHLEAF, CHUC, CHV2, CHB2, FPICE;
  // --- instrumentation the Fortran lacks
  float    ERRWAT;      // func:728 (computed, never asserted)
  float    ERRSW, ERRENG;
  uint32_t flags;
  int      iter_vege;   // loop1 passes executed (func:2744-2877)
  int      iter_stom_sun, iter_stom_sha;  // bisection passes (func:3832-3842)
  int      niter_soil;  // soilh2o substeps (func:5971-5978)
};

struct Ctx {
  Options opt;
  Tables  tbl;
};

// physics entry (driver.cpp)
void noahmp_sflx(const Ctx& cx, ColumnIO& c);

// leaf routines exposed for unit tests
void esat(float T, float& ESW, float& ESI, float& DESW, float& DESI);
void rosr12(float* P, const float* A, const float* B, float* C, const float* D, float* DELTA, int NTOP,
            int NSOIL_, int NSNOW_);
void combo(float& DZ, float& WLIQ, float& WICE, float& T, float DZ2, float WLIQ2, float WICE2, float T2);
void tdfcnd(const Ctx& cx, int sltyp, float& DF, float SMC, float soilwat);
void twostream(const Ctx& cx, int lutyp, int IB, int IC, float COSZ, float VAI, float FWET, float T,
               const A2& ALBGRD, const A2& ALBGRI, const A2& RHO, const A2& TAU, float fveg, int IST,
               A2& FAB, A2& FRE, A2& FTD, A2& FTI, float& GDIR, A2& FREV, A2& FREG, float& BGAP,
               float& WGAP);
void stomata(const Ctx& cx, int lutyp, float igs, float sfcprs, float sfctmp, float apar, float tv,
             float ea, float ei, float o2, float co2, float foln, float btran, float rb, float& rs,
             float& psn, int& niter_used);
void frh2o(const Ctx& cx, int sltyp, float& FREE, float TKELV, float SMC, float soilwat);

}  // namespace noahmp_oracle
