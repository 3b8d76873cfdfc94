// noahmp_b200 -- the step kernel: one thread advances one land column through nsteps time steps.
//
// Per step this is noahmp_sflx (core/module_noahmp_func.f90:66-476): atm (479-531), layer thickness,
// root-zone temperature, phenology (534-630), energy, water, carbon, error (633-732) and the closing
// fix-ups.  State is loaded from the structure-of-arrays planes once, kept in registers across all
// steps of the launch (time fusion: only the 9 forcing planes are read per step) and stored once.
#pragma once
#include "water.cuh"
#include "forcing_prep.cuh"

namespace nmp {

// One model step of a records-driven launch (include/noahmp_b200.h: noahmp_step_forcing): the forcing of the step is
// the time interpolation of two records and COSZ follows from the column's position - the arithmetic of
// forcing_prep.cuh, done in the load of the step kernel instead of in a pass of its own.
struct StepForcing {
  int rec0, rec1;           // the two bracketing records (slots of the record buffer)
  float w;                  // weight of rec1 (0: rec1 is not read)
  float sd, cd, ch0, sh0;   // solar scalars of the step
  int sp0, sp1;             // bit q set: plane q of rec0 / rec1 is zero-suppressed (ld_record); 0 for dense records
};
constexpr int MAX_RECORD_STEPS = 32;   // steps of one records-driven launch (the table travels as kernel parameters)

struct StepArgs {
  float* state;           // [NSTATE][ncol]
  const float* forcing;   // step s at forcing + s*forcing_stride, [NFORCING][ncol]; null in records mode
  int64_t forcing_stride;
  const float* records;   // records mode: record r at records + r*record_stride, [8][ncol] (+ [8][2][ceil(ncol/32)] words when packed)
  int64_t record_stride;
  const float* geo;       // records mode: [4][ncol] sin(lat) cos(lat) sin(lon) cos(lon)
  StepForcing tab[MAX_RECORD_STEPS];
  const float* static_f;  // [NSTATIC_F][ncol]
  const int* static_i;    // [ncol]
  float* diag;            // [NDIAG][ncol] or null
  double* budget;         // [NBUDGET_REP][NBUDGET] copies, all zero before the launch (budget_combine_kernel), or null
  const Tables* tables;   // global copy, staged to shared memory
  int64_t ncol;
  int nsteps;
  int final_launch;       // 1: add the end-of-launch budget slots (NCELL, SNEQV, SOILWATER)
  RunConst rc;
  OptRuntime opt;
};

// Forcing is read once: keep it out of L1, which holds the register spills of the column.
__device__ __forceinline__ float ld_forcing(const float* p) {
#ifdef NMP_PREFETCH_FORCING
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

// One plane of a forcing record.  Dense: element ii of plane q.  Zero-suppressed (noahmp_advance_records_host packs the
// planes that are zero over much of the grid - night-time SOLDN, dry PRCP - before they cross the host link): behind the
// eight planes the record carries, per plane, a 32-column occupancy word and the number of non-zero words before that
// group; the non-zero words themselves lie compacted at the start of the plane.  Lossless: a stored word is returned as it
// is, a suppressed one is +0 (the packer suppresses the bit pattern 0 only, so -0 travels).
// PACKED is a template parameter of the kernel so that launches on dense records run the code they always ran.
template <bool PACKED>
__device__ __forceinline__ float ld_record(const float* rec, int q, int sparse, int64_t n, int64_t ii) {
  if (!PACKED || !((sparse >> q) & 1)) return ld_forcing(rec + q * n + ii);
  const int64_t G = (n + 31) >> 5;
  const uint32_t* aux = reinterpret_cast<const uint32_t*>(rec + 8 * n) + (int64_t)(2 * q) * G + (ii >> 5);
  const uint32_t m = __ldg(aux), before = __ldg(aux + G);
  const unsigned lane = (unsigned)ii & 31u;
  float v = 0.0f;
  if ((m >> lane) & 1u) v = ld_forcing(rec + q * n + before + __popc(m & ((1u << lane) - 1u)));
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

constexpr int NBUDGET_REP = 64;   // copies of the budget vector the CTAs add into (one 128-byte line each)

// Folds the copies into the caller's vector (which accumulates over launches) and clears them for the next launch.
__global__ void budget_combine_kernel(double* __restrict__ rep, double* __restrict__ dst) {
  const int q = threadIdx.x;
  if (q >= NBUDGET) return;
  double r = 0.0;
  for (int k = 0; k < NBUDGET_REP; ++k) {
    const double x = rep[k * NBUDGET + q];
    r = (q == B_ERRWAT_MAXABS) ? fmax(r, x) : r + x;
    rep[k * NBUDGET + q] = 0.0;
  }
  dst[q] = (q == B_ERRWAT_MAXABS) ? fmax(dst[q], r) : dst[q] + r;
}

// Sums of up to 16 per-lane doubles over the warp, all at once: at every level of the butterfly a lane keeps half of the
// values it holds and hands the other half to its partner, so 8 + 4 + 2 + 1 + 1 exchanges do what 16 x 5 would (the
// shuffle pipe moves one warp-instruction per clock per SM: the plain form cost the budget builds 12 % of the step).
// Lane l returns the total of value l >> 1.
template <int N>
__device__ __forceinline__ double warp_sums(const double (&v)[N]) {
  static_assert(N <= 16, "one value per lane pair");
  const unsigned lane = threadIdx.x & 31u;
  double a8[8], a4[4], a2[2];
  static_for<0, 7>([&](auto j_) { constexpr int j = decltype(j_)::value;
    const double lo = (j < N) ? v[j < N ? j : 0] : 0.0, hi = (j + 8 < N) ? v[j + 8 < N ? j + 8 : 0] : 0.0;
    const bool up = (lane & 16u) != 0;
    a8[j] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, 16);
  });
  static_for<0, 3>([&](auto j_) { constexpr int j = decltype(j_)::value;
    const bool up = (lane & 8u) != 0;
    a4[j] = (up ? a8[j + 4] : a8[j]) + __shfl_xor_sync(0xffffffffu, up ? a8[j] : a8[j + 4], 8);
  });
  static_for<0, 1>([&](auto j_) { constexpr int j = decltype(j_)::value;
    const bool up = (lane & 4u) != 0;
    a2[j] = (up ? a4[j + 2] : a4[j]) + __shfl_xor_sync(0xffffffffu, up ? a4[j] : a4[j + 2], 4);
  });
  const bool up = (lane & 2u) != 0;
  double r = (up ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, up ? a2[0] : a2[1], 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// N budget slots starting at Q0, summed (B_ERRWAT_MAXABS: maximum) over the CTA: the warp (warp_sums), the warps through
// shared memory, then ONE atomic per CTA and slot into one of NBUDGET_REP copies of the vector.  Called by every thread of
// the CTA at the same place; consecutive calls alternate between two buffers (a warp may write the next call's partial
// sums while another still reads this call's: the barrier of the call in between orders the re-use).
constexpr int NBUDGET_STEP = B_RUNOFF_DT - B_ERRWAT_SUM + 1;   // the slots that sum over the steps
template <int Q0, int N>
__device__ __forceinline__ void budget_reduce(double* budget, double (*red)[NMP_BLOCK / 32], const double (&v)[N]) {
  constexpr int JMAX = B_ERRWAT_MAXABS - Q0;   // the slot that takes the maximum, if it is among these
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const double r = warp_sums<N>(v);            // (the maximum slot's sum is computed too and not used)
  if ((lane & 1u) == 0 && (lane >> 1) < (unsigned)N && (int)(lane >> 1) != JMAX) red[lane >> 1][warp] = r;
  if (JMAX >= 0 && JMAX < N) {
    const double m = warp_max(v[JMAX >= 0 && JMAX < N ? JMAX : 0]);
    if (lane == 0) red[JMAX >= 0 && JMAX < N ? JMAX : 0][warp] = m;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int j = threadIdx.x, q = Q0 + j, nw = (int)(blockDim.x >> 5);
    double t = red[j][0];
    if (q == B_ERRWAT_MAXABS) {
      for (int k = 1; k < nw; ++k) t = fmax(t, red[j][k]);
    } else {
      for (int k = 1; k < nw; ++k) t += red[j][k];
    }
    double* dst = budget + (size_t)(blockIdx.x & (NBUDGET_REP - 1)) * NBUDGET + q;
    if (q == B_ERRWAT_MAXABS) atomicMax(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)__double_as_longlong(t));
    else atomicAdd(dst, t);
  }
}

template <class Opt, bool EXT, bool DIAG, bool BUDGET, bool PACKED = false>
#ifndef NMP_MIN_BLOCKS
#define NMP_MIN_BLOCKS 2
#endif
#if NMP_MIN_BLOCKS > 0
__global__ void __launch_bounds__(NMP_BLOCK, NMP_MIN_BLOCKS) column_step_kernel(const StepArgs a) {
#else   // explicit register budget (tuning builds): -DNMP_MIN_BLOCKS=0 -DNMP_MAXNREG=n
__global__ void __maxnreg__(NMP_MAXNREG) column_step_kernel(const StepArgs a) {
#endif
  __shared__ Tables T;
  {  // stage the parameter tables (about 9.5 KB) into shared memory, coalesced
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.tables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
    for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < a.ncol;
  const int64_t ii = active ? i : a.ncol - 1;   // tail threads shadow the last column, never store
  const int64_t n = a.ncol;
  const Opt opt(a.opt);
  const RunConst& rc = a.rc;
  const float DT = rc.dt;

  // ---- load the column -------------------------------------------------------------------------
  Column c;
  {
    const float* S = a.state + ii;
    c.ALBOLD = S[S_ALBOLD * n]; c.SNEQVO = S[S_SNEQVO * n]; c.EAH = S[S_EAH * n]; c.FWET = S[S_FWET * n];
    c.CANLIQ = S[S_CANLIQ * n]; c.CANICE = S[S_CANICE * n]; c.TV = S[S_TV * n]; c.TG = S[S_TG * n];
    c.QSNOW = S[S_QSNOW * n]; c.ISNOW = __float_as_int(S[S_ISNOW * n]); c.snowh = S[S_SNOWH * n];
    c.sneqv = S[S_SNEQV * n]; c.ZWT = S[S_ZWT * n]; c.WA = S[S_WA * n]; c.WT = S[S_WT * n]; c.TAUSS = S[S_TAUSS * n];
    static_for<0, (7) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value; c.STC[k] = S[(S_STC + k) * n]; c.ZSNSO[k] = S[(S_ZSNSO + k) * n]; });
    static_for<0, (4) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value; c.soilwat[k] = S[(S_SOILWAT + k) * n]; c.SMC[k] = S[(S_SMC + k) * n]; });
    static_for<0, (3) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value; c.SNICE[k] = S[(S_SNICE + k) * n]; c.SNLIQ[k] = S[(S_SNLIQ + k) * n]; });
    if (EXT) {
      c.QSFC = S[S_QSFC * n]; c.TAH = S[S_TAH * n]; c.CM = S[S_CM * n]; c.CH = S[S_CH * n]; c.WSLAKE = S[S_WSLAKE * n];
      c.LFMASS = S[S_LFMASS * n]; c.RTMASS = S[S_RTMASS * n]; c.STMASS = S[S_STMASS * n]; c.WOOD = S[S_WOOD * n];
      c.STBLCP = S[S_STBLCP * n]; c.FASTCP = S[S_FASTCP * n]; c.LAI = S[S_LAI * n]; c.SAI = S[S_SAI * n];
    } else {  // dead under the shipped option set: overwritten before any read
      c.QSFC = 0.0f; c.TAH = 0.0f; c.CM = 0.0f; c.CH = 0.0f; c.WSLAKE = 0.0f; c.LFMASS = 0.0f; c.RTMASS = 0.0f;
      c.STMASS = 0.0f; c.WOOD = 0.0f; c.STBLCP = 0.0f; c.FASTCP = 0.0f; c.LAI = 0.0f; c.SAI = 0.0f;
    }
    c.LAT = a.static_f[C_LAT * n + ii];
    c.SHDFAC = a.static_f[C_SHDFAC * n + ii];
    c.TBOT = a.static_f[C_TBOT * n + ii];
    c.SHDMAX = (OptTraits<Opt>::runtime) ? a.static_f[C_SHDMAX * n + ii] : 0.0f;
    const int ty = a.static_i[ii];
    c.lutyp = ty & 0xff; c.sltyp = (ty >> 8) & 0xff; c.slptyp = (ty >> 16) & 0xff; c.ISC = (ty >> 24) & 0xf;
    c.IST = ((ty >> 28) & 1) + 1; c.ICE = (ty >> 29) & 1;
  }
  // A type index outside the tables would read past them (the reference, without bounds checks, would too): such a column
  // is flagged and left untouched; the arithmetic below runs on type 1 only so that the thread keeps step with its CTA.
  const bool badtype = c.lutyp < 1 || c.lutyp > T.nvegtyp || c.sltyp < 1 || c.sltyp > T.nsoiltyp || c.ISC < 1 ||
                       c.ISC > T.ncolor || c.slptyp < 1 || c.slptyp > NGEN - G_SLOPE;
  if (badtype) { c.lutyp = 1; c.sltyp = 1; c.slptyp = 1; c.ISC = 1; }
  const float* VG = &T.veg[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGF];
  const int ISURBAN = T.flags[0], ISWATER = T.flags[1], ISBARREN = T.flags[2], ISICE = T.flags[3];
  const int NROOT = T.vegi[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGI + VI_NROOT];

  DiagSink<DIAG> dg;
  // only the owner of a column writes its diagnostics: tail threads shadow the last column and a column with a type
  // outside the tables runs on substitute type 1 just to keep step with its CTA - neither may store (see below)
  dg.base = (DIAG && active && !badtype) ? a.diag + i : nullptr;
  dg.stride = n;
  unsigned anyflag = badtype ? (unsigned)FLAG_BADTYPE : 0u;
  float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // ACCPRCP ACCECAN ACCETRAN ACCEDIR SFCRNOFF UGDRNOFF (DIAG builds)
  if (DIAG && dg.base != nullptr) {
    static_for<0, 5>([&](auto q_) { constexpr int q = decltype(q_)::value; acc[q] = dg.base[(int64_t)(D_ACCPRCP + q) * n]; });
  }
  // budget sums: reduced over the CTA at the end of EVERY step (budget_reduce) - no fp64 accumulator lives across the step
  __shared__ double red[BUDGET ? 2 : 1][BUDGET ? NBUDGET_STEP : 1][BUDGET ? NMP_BLOCK / 32 : 1];

#pragma unroll 1
  for (int s = 0; s < a.nsteps; ++s) {
    Forcing f;
    if (a.records != nullptr) {   // records mode (uniform branch): interpolate + COSZ here, bit-identical to forcing_prep.cuh
      const StepForcing sf = a.tab[s];
      const float* R = a.records + (int64_t)sf.rec0 * a.record_stride;
      float v[8];
      if (sf.w == 0.0f) {
        static_for<0, 7>([&](auto q_) { constexpr int q = decltype(q_)::value; v[q] = ld_record<PACKED>(R, q, sf.sp0, n, ii); });
      } else {
        const float* R1 = a.records + (int64_t)sf.rec1 * a.record_stride;
        static_for<0, 7>([&](auto q_) { constexpr int q = decltype(q_)::value;
          v[q] = lerp_rn(ld_record<PACKED>(R, q, sf.sp0, n, ii), ld_record<PACKED>(R1, q, sf.sp1, n, ii), sf.w);
        });
      }
      f.SFCTMP = v[F_SFCTMP]; f.SFCPRS = v[F_SFCPRS]; f.UU = v[F_UU]; f.VV = v[F_VV]; f.Q2 = v[F_Q2]; f.SOLDN = v[F_SOLDN];
      f.LWDN = v[F_LWDN]; f.PRCP = v[F_PRCP];
      PrepScalars ps;
      ps.w = sf.w; ps.sd = sf.sd; ps.cd = sf.cd; ps.ch0 = sf.ch0; ps.sh0 = sf.sh0;
      const float* G = a.geo + ii;
      f.COSZ = cosz_rn(ld_forcing(G), ld_forcing(G + n), ld_forcing(G + 2 * n), ld_forcing(G + 3 * n), ps);
#ifdef NMP_PREFETCH_FORCING
      if (s + 1 < a.nsteps) {
        const float* Rn = a.records + (int64_t)a.tab[s + 1].rec0 * a.record_stride + ii;
        const int spn = PACKED ? a.tab[s + 1].sp0 : 0;   // a packed plane's words are not at ii: not prefetched
        static_for<0, 7>([&](auto q_) { constexpr int q = decltype(q_)::value;
          if (!((spn >> q) & 1)) asm volatile("prefetch.global.L2 [%0];" ::"l"(Rn + q * n));
        });
      }
#endif
    } else {
      const float* F = a.forcing + (int64_t)s * a.forcing_stride + ii;
      f.SFCTMP = ld_forcing(F + F_SFCTMP * n); f.SFCPRS = ld_forcing(F + F_SFCPRS * n); f.UU = ld_forcing(F + F_UU * n);
      f.VV = ld_forcing(F + F_VV * n); f.Q2 = ld_forcing(F + F_Q2 * n); f.SOLDN = ld_forcing(F + F_SOLDN * n);
      f.LWDN = ld_forcing(F + F_LWDN * n); f.PRCP = ld_forcing(F + F_PRCP * n); f.COSZ = ld_forcing(F + F_COSZ * n);
#ifdef NMP_PREFETCH_FORCING
      // the next step's nine forcing lines are asked for now (into L2, no register held), a whole step before the
      // CTA stops at its first barrier waiting for them
      if (s + 1 < a.nsteps) {
        const float* Fn = F + a.forcing_stride;
        static_for<0, NFORCING - 1>([&](auto q_) { constexpr int q = decltype(q_)::value;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(Fn + q * n));
        });
      }
#endif
    }
    const bool last = (s == a.nsteps - 1);
    DiagSink<DIAG> dgs = dg;
    if (DIAG && !last) dgs.base = nullptr;   // diagnostics of the last step only
    // driver duty: day of year, wrapped into [0, YEARLEN) so that a launch may cross 31 December
    // (phenology's month index IT1 = int(12 JULIAN / YEARLEN + 0.5) must stay in 0..12, func:586-594)
    const double jd = rc.julian0 + (double)s * (double)DT / 86400.0;
    const float JULIAN = (float)(jd - (double)rc.yearlen * floor(jd / (double)rc.yearlen));
    unsigned flags = 0;

  NMP_PHASE_SYNC_AT(0);
    // ---- atm, func:479-531 ------------------------------------------------------------------------
    const float THAIR = f.SFCTMP * m_powf(f.SFCPRS / f.SFCPRS, RAIR / CPAIR);
    const float QAIR = f.Q2;
    const float EAIR = QAIR * f.SFCPRS / (0.622f + 0.378f * QAIR);
    const float RHOAIR = (f.SFCPRS - 0.378f * EAIR) / (RAIR * f.SFCTMP);
    const float SWDOWN = (f.COSZ <= 0.0f) ? 0.0f : f.SOLDN;

    // ---- FICEOLD (driver duty in the reference) and layer thickness, func:322-328 ------
    float FICEOLD[3], DZSNSO[7];
    static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
      FICEOLD[IX3(k)] = (k >= c.ISNOW + 1) ? c.SNICE[IX3(k)] / (c.SNICE[IX3(k)] + c.SNLIQ[IX3(k)]) : 0.0f;
    });
    layer_thickness(c, DZSNSO);
    // ---- root-zone temperature and initial water storage, func:332-344 ---------------------------
    float TROOT = 0.0f;
    {
      const float zroot = root_depth(rc, NROOT);
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k <= NROOT) TROOT = TROOT + c.STC[IX7(k)] * DZSNSO[IX7(k)] / zroot;
      });
    }
    float BEG_WB = 0.0f;
    if (c.IST == 1) {
      BEG_WB = c.CANLIQ + c.CANICE + c.sneqv + c.WA;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        BEG_WB = BEG_WB + c.SMC[IX4(k)] * DZSNSO[IX7(k)] * 1000.0f;
      });
    }
    // ---- phenology, func:534-630 --------------------------------------------------------------------
    float elai, esai, IGS, HTOP;
    {
      float LAI = c.LAI, SAI = c.SAI;
      if (opt.veg == 1 || opt.veg == 3 || opt.veg == 4) {
        const float ylen = (float)rc.yearlen;
        const float DAY = (c.LAT >= 0.0f) ? JULIAN : fmodf(JULIAN + (0.5f * ylen), ylen);
        const float Tm = 12.0f * DAY / ylen;
        int IT1 = (int)(Tm + 0.5f);
        int IT2 = IT1 + 1;
        const float WT1 = ((float)IT1 + 0.5f) - Tm;
        const float WT2 = 1.0f - WT1;
        if (IT1 < 1) IT1 = 12;
        if (IT2 > 12) IT2 = 1;
        IT1 = NMP_CHK(IT1, 1, 12, 3);
        IT2 = NMP_CHK(IT2, 1, 12, 3);
        LAI = WT1 * VG[V_LAI12M + IT1 - 1] + WT2 * VG[V_LAI12M + IT2 - 1];
        SAI = WT1 * VG[V_SAI12M + IT1 - 1] + WT2 * VG[V_SAI12M + IT2 - 1];
      }
      if (SAI < 0.05f) SAI = 0.0f;
      if (LAI < 0.05f || SAI == 0.0f) LAI = 0.0f;
      if ((c.lutyp == ISWATER) || (c.lutyp == ISBARREN) || (c.lutyp == ISICE) || (c.lutyp == ISURBAN)) { LAI = 0.0f; SAI = 0.0f; }
      const float HVT = VG[V_HVT], HVB = VG[V_HVB];
      const float DB = fmin2(fmax2(c.snowh - HVB, 0.0f), HVT - HVB);
      float FB = DB / fmax2(1.0E-06f, HVT - HVB);
      if (HVT > 0.0f && HVT <= 1.0f) {
        const float snowhc = HVT * m_expf(-c.snowh / 0.2f);
        FB = fmin2(c.snowh, snowhc) / snowhc;
      }
      elai = LAI * (1.0f - FB);
      esai = SAI * (1.0f - FB);
      if (esai < 0.05f) esai = 0.0f;
      if (elai < 0.05f || esai == 0.0f) elai = 0.0f;
      IGS = (c.TV > VG[V_TMIN]) ? 1.0f : 0.0f;
      HTOP = HVT;
      c.LAI = LAI;
      c.SAI = SAI;
    }
    // ---- vegetation fraction, func:366-380 ----------------------------------------------------------
    float fveg = 0.0f;
    if (opt.veg == 1) { fveg = c.SHDFAC; if (fveg <= 0.01f) fveg = 0.01f; }
    else if (opt.veg == 2 || opt.veg == 3) { fveg = 1.0f - m_expf(-0.52f * (c.LAI + c.SAI)); if (fveg <= 0.01f) fveg = 0.01f; }
    else if (opt.veg == 4 || opt.veg == 5) { fveg = c.SHDMAX; if (fveg <= 0.01f) fveg = 0.01f; }
    else flags |= FLAG_OPTVEG;
    if (c.lutyp == ISURBAN || c.lutyp == ISBARREN) fveg = 0.0f;
    if (elai + esai == 0.0f) fveg = 0.0f;

    // ---- energy -------------------------------------------------------------------------------------
    EnergyResult e;
    energy<Opt, DIAG>(opt, T, rc, c, f, RHOAIR, QAIR, THAIR, EAIR, SWDOWN, IGS, HTOP, elai, esai, fveg, DZSNSO, e, dgs);
    flags |= e.flags;
    layer_thickness_again(c, DZSNSO);   // energy does not change the layering; not kept alive across it

    float soilice[4];
    static_for<0, (4) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value;
      soilice[k] = fmax2(0.0f, c.SMC[k] - c.soilwat[k]);
    });            // func:414
    c.SNEQVO = c.sneqv;                                                                        // func:415
    const float QVAP = fmax2(e.FGEV / e.LATHEAG, 0.0f);                                        // func:417-419
    const float QDEW = fabsf(fmin2(e.FGEV / e.LATHEAG, 0.0f));
    const float EDIR = QVAP - QDEW;

  NMP_PHASE_SYNC_AT(1);
    // ---- water ----------------------------------------------------------------------------------------
    WaterResult w;
    const float QPRECC = 0.10f * f.PRCP, QPRECL = 0.90f * f.PRCP;                                // func:517-518 (atm)
    water<Opt>(opt, T, rc, c, f, e, QPRECC, QPRECL, elai, esai, fveg, QVAP, QDEW, FICEOLD, soilice, DZSNSO, w);

  NMP_PHASE_SYNC_AT(2);
    // ---- carbon, func:439-447 ---------------------------------------------------------------------------
    float GPP = 0.0f, NPP = 0.0f, NEE = 0.0f;
    if (OptTraits<Opt>::runtime) {
      if (opt.veg == 2 || opt.veg == 5) carbon(T, rc, c, DZSNSO, e.PSN, e.BTRAN, IGS, GPP, NPP, NEE);
    }

    // ---- error, func:633-732: closure residuals become flag bits ---------------------------------------
    const float ERRSW = SWDOWN - (e.FSA + e.FSR);
    if (fabsf(ERRSW) > 0.01f) flags |= FLAG_ERRSW;
    const float ERRENG = e.SAV + e.SAG - (e.FIRA + e.FSH + e.FCEV + e.FGEV + e.FCTR + e.SSOIL);
    if (fabsf(ERRENG) > 0.01f) flags |= FLAG_ERRENG;
    float ERRWAT = 0.0f;
    if (c.IST == 1) {
      float END_WB = c.CANLIQ + c.CANICE + c.sneqv + c.WA;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        END_WB = END_WB + c.SMC[IX4(k)] * DZSNSO[IX7(k)] * 1000.0f;
      });
      ERRWAT = END_WB - BEG_WB - (f.PRCP - w.ECAN - w.ETRAN - EDIR - w.runsrf - w.runsub) * DT;
    }
    // ---- closing fix-ups, func:459-474 -------------------------------------------------------------------
    float Q2B = e.Q2B;
    if (c.lutyp == ISURBAN) {
      const float QFX = w.ETRAN + w.ECAN + EDIR;
      c.QSFC = (QFX / RHOAIR * e.CH_out) + QAIR;
      Q2B = c.QSFC;
    }
    if (c.snowh <= 1.0E-6f || c.sneqv <= 1.0E-3f) { c.snowh = 0.0f; c.sneqv = 0.0f; }
    const float ALBEDO = (SWDOWN != 0.0f) ? e.FSR / SWDOWN : -999.9f;

    {  // added instrumentation: non-finite prognostics
      bool bad = !isfinite(c.TV) || !isfinite(c.TG) || !isfinite(c.sneqv) || !isfinite(c.ZWT) || !isfinite(c.WA);
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        bad = bad || !isfinite(c.STC[IX7(k)]) || !isfinite(c.SMC[IX4(k)]);
      });
      if (bad) flags |= FLAG_NAN;
    }
    anyflag |= flags;

    if (DIAG) {
      dgs.put(D_ECAN, w.ECAN); dgs.put(D_ETRAN, w.ETRAN); dgs.put(D_EDIR, EDIR); dgs.put(D_Q2B, Q2B);
      dgs.put(D_RUNSRF, w.runsrf); dgs.put(D_RUNSUB, w.runsub); dgs.put(D_NEE, NEE); dgs.put(D_GPP, GPP);
      dgs.put(D_NPP, NPP); dgs.put(D_FVEG, fveg); dgs.put(D_ALBEDO, ALBEDO); dgs.put(D_QSNBOT, w.QSNBOT);
      dgs.put(D_PONDING1, w.PONDING1); dgs.put(D_PONDING2, w.PONDING2); dgs.put(D_FPICE, w.FPICE);
      dgs.put(D_ERRWAT, ERRWAT); dgs.put(D_ERRSW, ERRSW); dgs.put(D_ERRENG, ERRENG);
      dgs.puti(D_ITER, (e.iter_vege & 0xff) | ((e.iter_stom_sun & 0xff) << 8) | ((e.iter_stom_sha & 0xff) << 16) |
                           ((w.niter_soil & 0xff) << 24));
      dgs.puti(D_FLAGS, (int)anyflag);
    }
    if (DIAG) {   // accumulations between outputs (driver duty; offline/noahmp_config.py:83-84)
      acc[0] = acc[0] + f.PRCP * DT; acc[1] = acc[1] + w.ECAN * DT; acc[2] = acc[2] + w.ETRAN * DT;
      acc[3] = acc[3] + EDIR * DT; acc[4] = acc[4] + w.runsrf * DT; acc[5] = acc[5] + w.runsub * DT;
    }
    if constexpr (BUDGET) {
      // selects, not products: a flagged column's values may be NaN/Inf (0 x NaN = NaN would poison the domain sums)
      const bool on = active && !badtype;
      auto keep = [&](float x) { return (on && isfinite(x)) ? (double)x : 0.0; };
      double v[NBUDGET_STEP];   // slots B_ERRWAT_SUM .. B_RUNOFF_DT
      v[0] = keep(ERRWAT); v[1] = keep(fabsf(ERRWAT)); v[2] = v[1]; v[3] = keep(e.FSA); v[4] = keep(e.FSR); v[5] = keep(e.FIRA);
      v[6] = keep(e.FSH); v[7] = keep(e.FCEV + e.FGEV + e.FCTR); v[8] = keep(e.SSOIL); v[9] = keep(f.PRCP * DT);
      v[10] = keep((w.ECAN + w.ETRAN + EDIR) * DT); v[11] = keep((w.runsrf + w.runsub) * DT);
      budget_reduce<B_ERRWAT_SUM, NBUDGET_STEP>(a.budget, red[s & 1], v);
    }
  }

  if (DIAG && dg.base != nullptr) {
    static_for<0, 5>([&](auto q_) { constexpr int q = decltype(q_)::value; dg.base[(int64_t)(D_ACCPRCP + q) * n] = acc[q]; });
  }
  // a column whose type indices are outside the tables: state untouched, diagnostics zero, only its flag word is set
  if (DIAG && active && badtype && a.diag != nullptr) {
    for (int q = 0; q < NDIAG; ++q) a.diag[(int64_t)q * n + i] = (q == D_FLAGS) ? __int_as_float((int)anyflag) : 0.0f;
  }
  // ---- store the column -----------------------------------------------------------------------------
  if (active && !badtype) {
    float* S = a.state + i;
    S[S_ALBOLD * n] = c.ALBOLD; S[S_SNEQVO * n] = c.SNEQVO; S[S_EAH * n] = c.EAH; S[S_FWET * n] = c.FWET;
    S[S_CANLIQ * n] = c.CANLIQ; S[S_CANICE * n] = c.CANICE; S[S_TV * n] = c.TV; S[S_TG * n] = c.TG;
    S[S_QSNOW * n] = c.QSNOW; S[S_ISNOW * n] = __int_as_float(c.ISNOW); S[S_SNOWH * n] = c.snowh;
    S[S_SNEQV * n] = c.sneqv; S[S_ZWT * n] = c.ZWT; S[S_WA * n] = c.WA; S[S_WT * n] = c.WT; S[S_TAUSS * n] = c.TAUSS;
    static_for<0, (7) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value; S[(S_STC + k) * n] = c.STC[k]; S[(S_ZSNSO + k) * n] = c.ZSNSO[k]; });
    static_for<0, (4) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value; S[(S_SOILWAT + k) * n] = c.soilwat[k]; S[(S_SMC + k) * n] = c.SMC[k]; });
    static_for<0, (3) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value; S[(S_SNICE + k) * n] = c.SNICE[k]; S[(S_SNLIQ + k) * n] = c.SNLIQ[k]; });
    if (EXT) {
      S[S_QSFC * n] = c.QSFC; S[S_TAH * n] = c.TAH; S[S_CM * n] = c.CM; S[S_CH * n] = c.CH; S[S_WSLAKE * n] = c.WSLAKE;
      S[S_LFMASS * n] = c.LFMASS; S[S_RTMASS * n] = c.RTMASS; S[S_STMASS * n] = c.STMASS; S[S_WOOD * n] = c.WOOD;
      S[S_STBLCP * n] = c.STBLCP; S[S_FASTCP * n] = c.FASTCP; S[S_LAI * n] = c.LAI; S[S_SAI * n] = c.SAI;
    }
  }

  // ---- end-of-launch budget slots: cells, flagged columns, snow and soil water of the final state -------------------
  if constexpr (BUDGET) {
    const bool on = active && !badtype;
    const bool fin = active && a.final_launch;
    auto keep = [&](double x) { return (on && isfinite(x)) ? x : 0.0; };
    double soilw = 0.0;
    static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      const float thick = (k == 1) ? -rc.zsoil[0] : (rc.zsoil[k > 1 ? k - 2 : 0] - rc.zsoil[k - 1]);
      soilw += (double)(c.SMC[IX4(k)] * thick * 1000.0f);
    });
    double v2[2], v14[2];
    v2[0] = (fin && on) ? 1.0 : 0.0; v2[1] = (active && anyflag != 0) ? 1.0 : 0.0;
    v14[0] = fin ? keep((double)c.sneqv) : 0.0; v14[1] = fin ? keep(soilw) : 0.0;
    budget_reduce<B_NCELL, 2>(a.budget, red[a.nsteps & 1], v2);
    budget_reduce<B_SNEQV, 2>(a.budget, red[(a.nsteps + 1) & 1], v14);
  }
}

}  // namespace nmp
