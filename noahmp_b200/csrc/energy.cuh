// noahmp_b200 -- energy half of the column step, register-resident.
//
// One thread owns one land column.  Everything here is straight-line fp32 on scalars and small fixed
// arrays whose indices are compile-time constants (loops over layers are fully unrolled and guarded by
// `k >= ISNOW+1` predicates), so the column never leaves registers.  Each block cites the reference
// lines (core/module_noahmp_func.f90, "func:") whose arithmetic and operation order it reproduces.
#pragma once
#include "common.cuh"

namespace nmp {

// The column: prognostic state + per-column static data, all in registers.
struct Column {
  // static (SURVEY App. A group C)
  float LAT, SHDFAC, SHDMAX, TBOT;
  int lutyp, sltyp, slptyp, ISC, IST, ICE;
  // state (group S)
  float ALBOLD, SNEQVO, EAH, FWET, CANLIQ, CANICE, TV, TG, QSNOW, snowh, sneqv, ZWT, WA, WT, TAUSS;
  int ISNOW;
  float STC[7], soilwat[4], SMC[4], ZSNSO[7], SNICE[3], SNLIQ[3];
  // extended state (group s): live only for non-default options
  float QSFC, TAH, CM, CH, WSLAKE, LFMASS, RTMASS, STMASS, WOOD, STBLCP, FASTCP, LAI, SAI;
};

struct Forcing {
  float SFCTMP, SFCPRS, UU, VV, Q2, SOLDN, LWDN, PRCP, COSZ;
};

// What energy hands to water (the Fortran passes these through noahmp_sflx, func:388-433).
struct EnergyResult {
  float FSA, FSR, FIRA, FSH, SSOIL, FCEV, FGEV, FCTR, SAV, SAG, FSRV, FSRG;
  float BTRAN, BTRANI[4];
  float LATHEAG;
  float PONDING, QMELT;
  float APAR, PSN;
  float CH_out;            // tile-averaged CH (urban QSFC fix-up, func:461)
  float Q2B;
  int   IMELT_snow[3];     // IMELT(-2:0), the part water/compact read (func:4629,5589)
  bool  frozen_canopy, frozen_ground;
  unsigned flags;
  int   iter_vege, iter_stom_sun, iter_stom_sha;
};

// Diagnostics sink: stores go straight to the [NDIAG][ncol] planes when enabled, vanish otherwise.
template <bool ON>
struct DiagSink {
  float* base;     // diag + column index
  int64_t stride;  // ncol
  __device__ __forceinline__ void put(int plane, float v) const {
    if (ON && base != nullptr) base[(int64_t)plane * stride] = v;
  }
  __device__ __forceinline__ void puti(int plane, int v) const {
    if (ON && base != nullptr) base[(int64_t)plane * stride] = __int_as_float(v);
  }
};

// ---- func:3692-3736  saturation vapour pressure polynomials (Horner, as written) ------------------
__device__ __forceinline__ void esat(float T, float& ESW, float& ESI, float& DESW, float& DESI) {
  ESW = 100.0f * (6.107799961f + T * (4.436518521E-01f + T * (1.428945805E-02f + T * (2.650648471E-04f +
        T * (3.031240396E-06f + T * (2.034080948E-08f + T * 6.136820929E-11f))))));
  ESI = 100.0f * (6.109177956f + T * (5.034698970E-01f + T * (1.886013408E-02f + T * (4.176223716E-04f +
        T * (5.824720280E-06f + T * (4.838803174E-08f + T * 1.838826904E-10f))))));
  DESW = 100.0f * (4.438099984E-01f + T * (2.857002636E-02f + T * (7.938054040E-04f + T * (1.215215065E-05f +
         T * (1.036561403E-07f + T * (3.532421810e-10f + T * -7.090244804E-13f))))));
  DESI = 100.0f * (5.030305237E-01f + T * (3.773255020E-02f + T * (1.267995369E-03f + T * (2.477563108E-05f +
         T * (3.005693132E-07f + T * (2.158542548E-09f + T * 7.131097725E-12f))))));
}
__device__ __forceinline__ float tdc(float T) { return fmin2(50.0f, fmax2(-50.0f, (T - TFRZ))); }  // func:2688
// Every caller keeps the over-water pair for T > 0 and the over-ice pair otherwise (func:2789-2797, 2889-2897, 3176-3184,
// 2712-2718, 3225-3231).  Evaluating only the pair a warp needs (vote + branch) was measured and is SLOWER (2.51e9 ->
// 2.31e9 cell-steps/s): the 24 independent FMAs are cheap filler for the issue slots, the vote and branches are not.
__device__ __forceinline__ void esat_sel(float T, float& ES, float& DES) {
  float ESW, ESI, DESW, DESI;
  esat(T, ESW, ESI, DESW, DESI);
  if (T > 0.0f) { ES = ESW; DES = DESW; } else { ES = ESI; DES = DESI; }
}

// ---- func:1500-1595  Johansen / Peters-Lidard soil conductivity ------------------------------------
__device__ __forceinline__ float tdfcnd(float smcmax, float quartz, float SMC, float soilwat) {
  const float SATRATIO = SMC / smcmax;
  const float THKS = m_powf(7.7f, quartz) * m_powf(2.0f, 1.0f - quartz);                      // func:1561
  const float XU = (soilwat / SMC) * smcmax;                                              // func:1564-1566
  const float THKSAT = m_powf(THKS, 1.0f - smcmax) * m_powf(TKICE, smcmax - XU) * m_powf(0.57f, XU);  // func:1569
  const float GAMMD = (1.0f - smcmax) * 2700.0f;
  const float THKDRY = (0.135f * GAMMD + 64.7f) / (2700.0f - 0.947f * GAMMD);             // func:1574
  float AKE;
  if ((soilwat + 0.0005f) < SMC) AKE = SATRATIO;                                          // func:1576-1591
  else AKE = (SATRATIO > 0.1f) ? m_log10f(SATRATIO) + 1.0f : 0.0f;
  return AKE * (THKSAT - THKDRY) + THKDRY;
}

// ---- func:3353-3508  Monin-Obukhov exchange coefficients -------------------------------------------
// The four logs at func:3417-3420 depend only on heights, so the caller computes them once per Newton
// loop (identical values every pass) and hands them in.
struct MoninObukhov {
  float MOZ, FM, FH, FM2, FH2, FV, CM, CH, CH2;
  int MOZSGN;
};
__device__ __forceinline__ void sfcdif1(int iter, float SFCTMP, float RHOAIR, float H, float QAIR, float ZLVL,
                                        float ZPD, float Z0H, float UR, float TMPCM, float TMPCH, float TMPCM2,
                                        float TMPCH2, MoninObukhov& m) {
  const float MOZOLD = m.MOZ;
  float MOZ2;
  if (iter == 1) {                                                                        // func:3422-3434
    m.FV = 0.0f;
    m.MOZ = 0.0f;
    MOZ2 = 0.0f;
  } else {
    const float TVIR = (1.0f + 0.61f * QAIR) * SFCTMP;
#ifdef NMP_FAST_ALGEBRA   // one reciprocal instead of the chain /TVIR, /(RHOAIR CPAIR), /TMP1, /MOL
    float TMP1 = H * ((KARMAN * GRAV) / (TVIR * (RHOAIR * CPAIR)));
    if (fabsf(TMP1) <= MPE) TMP1 = MPE;
    const float RMOL = -TMP1 / cube(m.FV);
    m.MOZ = fmin2((ZLVL - ZPD) * RMOL, 1.0f);
    MOZ2 = fmin2((2.0f + Z0H) * RMOL, 1.0f);
#else
    float TMP1 = KARMAN * (GRAV / TVIR) * H / (RHOAIR * CPAIR);
    if (fabsf(TMP1) <= MPE) TMP1 = MPE;
    const float MOL = -1.0f * cube(m.FV) / TMP1;
    m.MOZ = fmin2((ZLVL - ZPD) / MOL, 1.0f);
    MOZ2 = fmin2((2.0f + Z0H) / MOL, 1.0f);
#endif
  }
  if (MOZOLD * m.MOZ < 0.0f) m.MOZSGN = m.MOZSGN + 1;                                     // func:3438-3446
  if (m.MOZSGN >= 2) {
    m.MOZ = 0.0f; m.FM = 0.0f; m.FH = 0.0f; MOZ2 = 0.0f; m.FM2 = 0.0f; m.FH2 = 0.0f;
  }
  float FMNEW, FHNEW, FM2NEW, FH2NEW;
  if (m.MOZ < 0.0f) {                                                                     // func:3449-3467
    const float T1 = m_powf(1.0f - 16.0f * m.MOZ, 0.25f);
    const float T2 = m_logf((1.0f + T1 * T1) / 2.0f);
    const float T3 = m_logf((1.0f + T1) / 2.0f);
    FMNEW = 2.0f * T3 + T2 - 2.0f * m_atanf(T1) + 1.5707963f;
    FHNEW = 2.0f * T2;
    const float T12 = m_powf(1.0f - 16.0f * MOZ2, 0.25f);
    const float T22 = m_logf((1.0f + T12 * T12) / 2.0f);
    const float T32 = m_logf((1.0f + T12) / 2.0f);
    FM2NEW = 2.0f * T32 + T22 - 2.0f * m_atanf(T12) + 1.5707963f;
    FH2NEW = 2.0f * T22;
  } else {
    FMNEW = -5.0f * m.MOZ;
    FHNEW = FMNEW;
    FM2NEW = -5.0f * MOZ2;
    FH2NEW = FM2NEW;
  }
  if (iter == 1) {                                                                        // func:3472-3482
    m.FM = FMNEW; m.FH = FHNEW; m.FM2 = FM2NEW; m.FH2 = FH2NEW;
  } else {
    m.FM = 0.5f * (m.FM + FMNEW);
    m.FH = 0.5f * (m.FH + FHNEW);
    m.FM2 = 0.5f * (m.FM2 + FM2NEW);
    m.FH2 = 0.5f * (m.FH2 + FH2NEW);
  }
  m.FH = fmin2(m.FH, 0.9f * TMPCH);                                                       // func:3486-3489
  m.FM = fmin2(m.FM, 0.9f * TMPCM);
  m.FH2 = fmin2(m.FH2, 0.9f * TMPCH2);
  m.FM2 = fmin2(m.FM2, 0.9f * TMPCM2);
  float CMFM = TMPCM - m.FM, CHFH = TMPCH - m.FH, CM2FM2 = TMPCM2 - m.FM2, CH2FH2 = TMPCH2 - m.FH2;  // func:3491-3498
  if (fabsf(CMFM) <= MPE) CMFM = MPE;
  if (fabsf(CHFH) <= MPE) CHFH = MPE;
  if (fabsf(CM2FM2) <= MPE) CM2FM2 = MPE;
  if (fabsf(CH2FH2) <= MPE) CH2FH2 = MPE;
#ifdef NMP_FAST_ALGEBRA   // sqrt(K^2 / CMFM^2) = K / |CMFM|: two independent reciprocals, no square root
  const float RCMFM = 1.0f / CMFM;
  m.CM = (KARMAN * KARMAN) * RCMFM * RCMFM;
  m.CH = (KARMAN * KARMAN) * RCMFM / CHFH;
  m.FV = UR * (KARMAN * fabsf(RCMFM));
#else
  m.CM = KARMAN * KARMAN / (CMFM * CMFM);                                                 // func:3499-3506
  m.CH = KARMAN * KARMAN / (CMFM * CHFH);
  m.FV = UR * sqrtf(m.CM);
#endif
  m.CH2 = KARMAN * m.FV / CH2FH2;
}

// ---- func:3511-3689  Chen97 exchange coefficients (opt_sfc = 2) ------------------------------------
__device__ __forceinline__ void sfcdif2(int iter, float Z0, float THZ0, float THLM, float SFCSPD, float CZIL,
                                     float ZLM, float& AKMS, float& AKHS, float& RLMO, float& WSTAR2,
                                     float& USTAR) {
  const float WWST2 = 1.2f * 1.2f, VKRM = 0.40f, EXCM = 0.001f, BTG = (1.0f / 270.0f) * GRAV, ELFC = VKRM * BTG;
  const float WOLD = 0.15f, WNEW = 1.0f - WOLD, PIHF = 3.14159265f / 2.0f, EPSU2 = 1.E-4f, EPSUST = 0.07f;
  const float ZTMIN = -5.0f, ZTMAX = 1.0f, HPBL = 1000.0f, SQVISC = 258.2f;
  auto PSPMU = [&](float XX) {
    return -2.0f * m_logf((XX + 1.0f) * 0.5f) - m_logf((XX * XX + 1.0f) * 0.5f) + 2.0f * m_atanf(XX) - PIHF;
  };
  auto PSPHU = [&](float XX) { return -2.0f * m_logf((XX * XX + 1.0f) * 0.5f); };
  const float ZILFC = -CZIL * VKRM * SQVISC;
  const float ZU = Z0, CXCH = EXCM * (1.0f / ZLM), DTHV = THLM - THZ0;
  const float DU2 = fmax2(SFCSPD * SFCSPD, EPSU2), BTGH = BTG * HPBL;
  if (iter == 1) {                                                                        // func:3599-3607
    if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * m_powf(fabsf(BTGH * AKHS * DTHV), 2.0f / 3.0f);
    else WSTAR2 = 0.0f;
    USTAR = fmax2(sqrtf(AKMS * sqrtf(DU2 + WSTAR2)), EPSUST);
    RLMO = ELFC * AKHS * DTHV / cube(USTAR);
  }
  float ZT = fmax2(1.0E-6f, m_expf(ZILFC * sqrtf(USTAR * Z0)) * Z0);                        // func:3610-3621
  const float ZSLU = ZLM + ZU;
  float ZSLT = ZLM + ZT;
  const float RLOGU = m_logf(ZSLU / ZU), RLOGT = m_logf(ZSLT / ZT);
  float ZETALT = fmax2(ZSLT * RLMO, ZTMIN);
  RLMO = ZETALT / ZSLT;
  float ZETALU = ZSLU * RLMO;
  const float ZETAU = ZU * RLMO, ZETAT = ZT * RLMO;
  float SIMM, SIMH;
  if (RLMO < 0.0f) {                                                                      // func:3624-3645 (ILECH = 0)
    const float XLU = sqrtf(sqrtf(1.0f - 16.0f * ZETALU)), XLT = sqrtf(sqrtf(1.0f - 16.0f * ZETALT));
    const float XU = sqrtf(sqrtf(1.0f - 16.0f * ZETAU)), XT = sqrtf(sqrtf(1.0f - 16.0f * ZETAT));
    const float PSMZ = PSPMU(XU);
    SIMM = PSPMU(XLU) - PSMZ + RLOGU;
    const float PSHZ = PSPHU(XT);
    SIMH = PSPHU(XLT) - PSHZ + RLOGT;
  } else {
    ZETALU = fmin2(ZETALU, ZTMAX);
    ZETALT = fmin2(ZETALT, ZTMAX);
    const float PSMZ = 5.0f * ZETAU;
    SIMM = 5.0f * ZETALU - PSMZ + RLOGU;
    const float PSHZ = 5.0f * ZETAT;
    SIMH = 5.0f * ZETALT - PSHZ + RLOGT;
  }
  USTAR = fmax2(sqrtf(AKMS * sqrtf(DU2 + WSTAR2)), EPSUST);                               // func:3664-3687
  const float USTARK = USTAR * VKRM;
  AKMS = fmax2(USTARK / SIMM, CXCH);
  AKHS = fmax2(USTARK / SIMH, CXCH);
  if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * m_powf(fabsf(BTGH * AKHS * DTHV), 2.0f / 3.0f);
  else WSTAR2 = 0.0f;
  const float RLMN = ELFC * AKHS * DTHV / cube(USTAR);
  RLMO = RLMO * WOLD + RLMN * WNEW;
}

// ---- func:3739-3887  Ball-Berry stomatal resistance, bisection on ci -------------------------------
// Written as a stepper (begin / step / rs) so the caller can run the bisection as a CTA-uniform loop with a
// barrier per pass; the arithmetic is that of stomata + its internal ci2ci (func:3847-3886).
struct LeafParams {  // PHOTO row of one land-use type
  float BP, MP, FOLNMX, QE25, KC25, AKC, KO25, AKO, VCMX25, AVCMX;
  int C3C4;
};
struct Stomata {
  float cf, rs, psn, j, awc, cp, vcmx, rlb, cihigh, cilow, wj, wc, we;
  int niter;
  bool done;
  __device__ __forceinline__ void begin(const LeafParams& L, float sfcprs, float sfctmp, float apar, float tv, float o2,
                                        float co2, float foln, float btran, float rb) {
    niter = 0;
    cf = sfcprs / (RGAS * sfctmp) * 1.0e06f;                                              // func:3801-3803
    rs = 1.0f / L.BP * cf;
    psn = 0.0f;
    done = false;
    wj = 0.0f; wc = 0.0f; we = 0.0f;
    j = 0.0f; awc = 0.0f; cp = 0.0f; vcmx = 0.0f; rlb = 0.0f; cihigh = 0.0f; cilow = 0.0f;
    if (apar <= 0.0f) { done = true; return; }                                            // func:3806 (rs is NOT rescaled)
    const float fnf = fmin2(foln / fmax2(MPE, L.FOLNMX), 1.0f);
    const float tc = tv - TFRZ;
    j = (4.6f * apar) * L.QE25;                                                           // func:3810-3811
    const float q10 = (tc - 25.0f) / 10.0f;
    const float kc = L.KC25 * m_powf(L.AKC, q10);                                         // func:3815-3818
    const float ko = L.KO25 * m_powf(L.AKO, q10);
    awc = kc * (1.0f + o2 / ko);
    cp = 0.5f * kc / ko * o2 * 0.21f;
    vcmx = L.VCMX25 / (1.0f + m_expf((-2.2E05f + 710.0f * (tc + TFRZ)) / (8.314f * (tc + TFRZ)))) * fnf * btran *
           (m_powf(L.AVCMX, q10));                                                        // func:3821-3824
    rlb = rb / cf;                                                                        // func:3827
    cihigh = 1.5f * co2;                                                                  // func:3830-3831
    cilow = 0.0f;
  }
  // one pass of the do-loop at func:3832-3842; sets done when the loop would exit
  __device__ __forceinline__ void step(const LeafParams& L, float igs, float sfcprs, float ea, float ei, float co2) {
    niter = niter + 1;
    const float ci = 0.5f * (cihigh + cilow);
    if (L.C3C4 == 1) {
      wj = fmax2(ci - cp, 0.0f) * j / (ci + 2.0f * cp);
      wc = fmax2(ci - cp, 0.0f) * vcmx / (ci + awc);
      we = 0.5f * vcmx;
    } else if (L.C3C4 == 2) {
      wj = j;
      wc = vcmx;
      we = 4000.0f * vcmx * ci / sfcprs;
    }
    psn = fmin2(fmin2(wj, wc), we) * igs;
    const float cs = fmax2(co2 - 1.37f * rlb * sfcprs * psn, MPE);
#ifdef NMP_FAST_ALGEBRA   // shared reciprocal of cs, one reciprocal for the two roots
    const float mpc = L.MP * psn * sfcprs / cs;
    const float a = mpc * (ea / ei) + L.BP;
    const float b = (mpc + L.BP) * rlb - 1.0f;
    const float c = -rlb;
    const float disc = sqrtf(b * b - 4.0f * a * c);
    const float q = (b >= 0.0f) ? -0.5f * (b + disc) : -0.5f * (b - disc);
    const float raq = 1.0f / (a * q);
    rs = fmax2(q * q * raq, c * a * raq);
#else
    const float a = L.MP * psn * sfcprs * ea / (cs * ei) + L.BP;
    const float b = (L.MP * psn * sfcprs / cs + L.BP) * rlb - 1.0f;
    const float c = -rlb;
    const float disc = sqrtf(b * b - 4.0f * a * c);
    const float q = (b >= 0.0f) ? -0.5f * (b + disc) : -0.5f * (b - disc);
    rs = fmax2(q / a, c / q);
#endif
    const float fci = fmax2(cs - psn * sfcprs * 1.65f * rs, 0.0f);
    if (((cihigh - cilow) <= 5.0E-2f) || fabsf(fci - ci) <= MPE) done = true;
    else if (fci > ci) cilow = ci;
    else cihigh = ci;
    if (niter >= 20) done = true;
    if (done) rs = rs * cf;                                                               // func:3844
  }
};

// ---- func:3890-3984  Jarvis canopy resistance (opt_crs = 2) ----------------------------------------
__device__ __forceinline__ void canres(float RGL, float RSMIN, float RSMAX, float TOPT, float HS, float sfcprs,
                                       float tv, float par, float eah, float btran, float& rs, float& psn) {
  float q2 = 0.622f * eah / (sfcprs - 0.378f * eah);
  q2 = q2 / (1.0f + q2);
  // calhum, func:3958-3984
  const float ES = 0.611f * m_expf(2.501E6f / 461.0f * (1.0f / 273.15f - 1.0f / tv));
  float q2sat = 0.622f * ES / (sfcprs * 1.E-3f - ES);
  q2sat = q2sat * 1.E3f;
  q2sat = q2sat / 1.E3f;
  const float ff = 2.0f * par / RGL;
  float rcs = (ff + RSMIN / RSMAX) / (1.0f + ff);
  rcs = fmin2(fmax2(rcs, 0.0001f), 1.0f);
  float rct = 1.0f - 0.0016f * (sq(TOPT - tv));
  rct = fmin2(fmax2(rct, 0.0001f), 1.0f);
  float rcq = 1.0f / (1.0f + HS * fmax2(0.0f, q2sat - q2));
  rcq = fmin2(fmax2(rcq, 0.01f), 1.0f);
  rs = RSMIN / (rcs * rct * rcq * btran);
  psn = __int_as_float(0x7fc00000);                                                       // func:3954 psn = NaN
}

// ---- func:2215-2462  two-stream canopy radiative transfer ------------------------------------------
// Quantities that do not depend on (band, direct/diffuse) are computed once by the caller (TwoStreamCommon).
struct TwoStreamCommon {
  float GAP, KOPEN, BGAP, WGAP, GDIR, EXT, AVMU, CHILfac /*((1+CHIL)/2)**2*/, PHI1, PHI2, COSZI;
};
__device__ __forceinline__ void twostream_common(int opt_rad, float COSZ, float VAI, float fveg, float XL,
                                                 float RCROWN, float HVT, float HVB, float BGAP_in,
                                                 float WGAP_in, TwoStreamCommon& t) {
  const float PAI = 3.14159265f;
  t.BGAP = BGAP_in;
  t.WGAP = WGAP_in;
  t.GAP = 0.0f;
  t.KOPEN = 0.0f;
  if (VAI == 0.0f) {                                                                      // func:2307-2309
    t.GAP = 1.0f;
    t.KOPEN = 1.0f;
  } else {
    if (opt_rad == 1) {                                                                   // func:2311-2324
      const float DENfveg = -m_logf(fmax2(1.0f - fveg, 0.01f)) / (PAI * sq(RCROWN));
      const float HD = HVT - HVB;
      const float BB = 0.5f * HD;
      const float THETAP = m_atanf(BB / RCROWN * m_tanf(m_acosf(fmax2(0.01f, COSZ))));
      t.BGAP = m_expf(-DENfveg * PAI * sq(RCROWN) / m_cosf(THETAP));
      const float FA = VAI / (1.33f * PAI * m_powf(RCROWN, 3.0f) * (BB / RCROWN) * DENfveg);
      const float NEWVAI = HD * FA;
      t.WGAP = (1.0f - t.BGAP) * m_expf(-0.5f * NEWVAI / COSZ);
      t.GAP = fmin2(1.0f - fveg, t.BGAP + t.WGAP);
      t.KOPEN = 0.05f;
    }
    if (opt_rad == 2) { t.GAP = 0.0f; t.KOPEN = 0.0f; }                                   // func:2326-2329
    if (opt_rad == 3) { t.GAP = 1.0f - fveg; t.KOPEN = 1.0f - fveg; }                     // func:2331-2334
  }
  t.COSZI = fmax2(0.001f, COSZ);                                                          // func:2344-2351
  float CHIL = fmin2(fmax2(XL, -0.4f), 0.6f);
  if (fabsf(CHIL) <= 0.01f) CHIL = 0.01f;
  t.PHI1 = 0.5f - 0.633f * CHIL - 0.330f * CHIL * CHIL;
  t.PHI2 = 0.877f * (1.0f - 2.0f * t.PHI1);
  t.GDIR = t.PHI1 + t.PHI2 * t.COSZI;
  t.EXT = t.GDIR / t.COSZI;
  t.AVMU = (1.0f - t.PHI1 / t.PHI2 * m_logf((t.PHI1 + t.PHI2) / t.PHI1)) / t.PHI2;
  t.CHILfac = sq((1.0f + CHIL) / 2.0f);
}

struct TwoStreamOut { float FAB, FRE, FTD, FTI, FREV, FREG; };
__device__ __forceinline__ void twostream_band(const TwoStreamCommon& t, int IC, float VAI, float FWET, float T,
                                               float ALBGRD, float ALBGRI, float RHO, float TAU, float OMEGAS,
                                               float BETADS, float BETAIS, TwoStreamOut& o) {
  const float OMEGAL = RHO + TAU;                                                         // func:2352-2358
  float TMP0 = t.GDIR + t.PHI2 * t.COSZI;
  float TMP1 = t.PHI1 * t.COSZI;
  const float ASU = 0.5f * OMEGAL * t.GDIR / TMP0 * (1.0f - TMP1 / TMP0 * m_logf((TMP1 + TMP0) / TMP1));
  const float BETADL = (1.0f + t.AVMU * t.EXT) / (OMEGAL * t.AVMU * t.EXT) * ASU;
  const float BETAIL = 0.5f * (RHO + TAU + (RHO - TAU) * t.CHILfac) / OMEGAL;
  float OMEGA, BETAD, BETAI;
  if (T > TFRZ) {                                                                         // func:2362-2374
    OMEGA = OMEGAL; BETAD = BETADL; BETAI = BETAIL;
  } else {
    OMEGA = (1.0f - FWET) * OMEGAL + FWET * OMEGAS;
    BETAD = ((1.0f - FWET) * OMEGAL * BETADL + FWET * OMEGAS * BETADS) / OMEGA;
    BETAI = ((1.0f - FWET) * OMEGAL * BETAIL + FWET * OMEGAS * BETAIS) / OMEGA;
  }
  const float B = 1.0f - OMEGA + OMEGA * BETAI;                                           // func:2378-2392
  const float C = OMEGA * BETAI;
  TMP0 = t.AVMU * t.EXT;
  const float D = TMP0 * OMEGA * BETAD;
  const float F = TMP0 * OMEGA * (1.0f - BETAD);
  TMP1 = B * B - C * C;
  const float H = sqrtf(TMP1) / t.AVMU;
  float SIGMA = TMP0 * TMP0 - TMP1;
  if (fabsf(SIGMA) < 1.e-6f) SIGMA = copysignf(1.e-6f, SIGMA);
#ifdef NMP_FAST_ALGEBRA
  const float AH = t.AVMU * H;
  const float P1 = B + AH, P2 = B - AH, P3 = B + TMP0, P4 = B - TMP0;
  const float S1 = m_expf(-H * VAI);
  const float S2 = m_expf(-t.EXT * VAI);
  const float ALB = (IC == 0) ? ALBGRD : ALBGRI;                                          // func:2393-2401
  const float U1 = B - C / ALB, U2 = B - C * ALB, U3 = F + C * ALB;
  const float TMP2 = U1 - AH, TMP3 = U1 + AH;                                             // func:2402-2421
  // the fast build divides once by each of S1, D1, D2, SIGMA (18 quotients in the reference's form)
  const float RS1 = 1.0f / S1, RSIG = 1.0f / SIGMA;
  const float D1 = P1 * TMP2 * RS1 - P2 * TMP3 * S1;
  const float TMP4 = U2 + AH, TMP5 = U2 - AH;
  const float D2 = TMP4 * RS1 - TMP5 * S1;
  const float RD1 = 1.0f / D1, RD2 = 1.0f / D2;
  const float H1 = -D * P4 - C * F;
  const float H1S = H1 * RSIG;
  const float TMP6 = D - H1S * P3;
  const float TMP7 = (D - C - H1S * (U1 + TMP0)) * S2;
  const float H2 = (TMP6 * TMP2 * RS1 - P2 * TMP7) * RD1;
  const float H3 = -(TMP6 * TMP3 * S1 - P1 * TMP7) * RD1;
  const float H4 = -F * P3 - C * D;
  const float TMP8 = H4 * RSIG;
  const float TMP9 = (U3 - TMP8 * (U2 - TMP0)) * S2;
  const float H5 = -(TMP8 * TMP4 * RS1 + TMP9) * RD2;
  const float H6 = (TMP8 * TMP5 * S1 + TMP9) * RD2;
  const float H7 = (C * TMP2) * RD1 * RS1;
  const float H8 = (-C * TMP3 * S1) * RD1;
  const float H9 = TMP4 * RD2 * RS1;
  const float H10 = (-TMP5 * S1) * RD2;
  if (IC == 0) {                                                                          // func:2426-2450
    o.FTD = S2 * (1.0f - t.GAP) + t.GAP;
    o.FTI = (TMP8 * S2 + H5 * S1 + H6 * RS1) * (1.0f - t.GAP);
    o.FRE = (H1S + H2 + H3) * (1.0f - t.GAP) + ALBGRD * t.GAP;
    o.FREV = (H1S + H2 + H3) * (1.0f - t.GAP);
    o.FREG = ALBGRD * t.GAP;
  } else {
    o.FTD = 0.0f;
    o.FTI = (H9 * S1 + H10 * RS1) * (1.0f - t.KOPEN) + t.KOPEN;
    o.FRE = (H7 + H8) * (1.0f - t.KOPEN) + ALBGRI * t.KOPEN;
    o.FREV = (H7 + H8) * (1.0f - t.KOPEN) + ALBGRI * t.KOPEN;
    o.FREG = 0.0f;
  }
#else
  const float AH = t.AVMU * H;
  const float P1 = B + AH, P2 = B - AH, P3 = B + TMP0, P4 = B - TMP0;
  const float S1 = m_expf(-H * VAI);
  const float S2 = m_expf(-t.EXT * VAI);
  const float ALB = (IC == 0) ? ALBGRD : ALBGRI;                                          // func:2393-2401
  const float U1 = B - C / ALB, U2 = B - C * ALB, U3 = F + C * ALB;
  const float TMP2 = U1 - AH, TMP3 = U1 + AH;                                             // func:2402-2421
  const float D1 = P1 * TMP2 / S1 - P2 * TMP3 * S1;
  const float TMP4 = U2 + AH, TMP5 = U2 - AH;
  const float D2 = TMP4 / S1 - TMP5 * S1;
  const float H1 = -D * P4 - C * F;
  const float TMP6 = D - H1 * P3 / SIGMA;
  const float TMP7 = (D - C - H1 / SIGMA * (U1 + TMP0)) * S2;
  const float H2 = (TMP6 * TMP2 / S1 - P2 * TMP7) / D1;
  const float H3 = -(TMP6 * TMP3 * S1 - P1 * TMP7) / D1;
  const float H4 = -F * P3 - C * D;
  const float TMP8 = H4 / SIGMA;
  const float TMP9 = (U3 - TMP8 * (U2 - TMP0)) * S2;
  const float H5 = -(TMP8 * TMP4 / S1 + TMP9) / D2;
  const float H6 = (TMP8 * TMP5 * S1 + TMP9) / D2;
  const float H7 = (C * TMP2) / (D1 * S1);
  const float H8 = (-C * TMP3 * S1) / D1;
  const float H9 = TMP4 / (D2 * S1);
  const float H10 = (-TMP5 * S1) / D2;
  if (IC == 0) {                                                                          // func:2426-2450
    o.FTD = S2 * (1.0f - t.GAP) + t.GAP;
    o.FTI = (H4 * S2 / SIGMA + H5 * S1 + H6 / S1) * (1.0f - t.GAP);
    o.FRE = (H1 / SIGMA + H2 + H3) * (1.0f - t.GAP) + ALBGRD * t.GAP;
    o.FREV = (H1 / SIGMA + H2 + H3) * (1.0f - t.GAP);
    o.FREG = ALBGRD * t.GAP;
  } else {
    o.FTD = 0.0f;
    o.FTI = (H9 * S1 + H10 / S1) * (1.0f - t.KOPEN) + t.KOPEN;
    o.FRE = (H7 + H8) * (1.0f - t.KOPEN) + ALBGRI * t.KOPEN;
    o.FREV = (H7 + H8) * (1.0f - t.KOPEN) + ALBGRI * t.KOPEN;
    o.FREG = 0.0f;
  }
#endif
  o.FAB = 1.0f - o.FRE - (1.0f - ALBGRD) * o.FTD - (1.0f - ALBGRI) * o.FTI;               // func:2454-2455
}

// ---- the canopy Newton loop ("loop1", func:2744-2877) as a relocatable object ------------------------
// Everything one pass of the loop reads or writes for one column, so a column's iteration can be handed to
// another thread of the CTA (tail compaction, see energy()).  first() = sfcdif1|2 + ragrb (func:3260-3350) +
// esat(TV); second() = canopy-air closure, fluxes, Newton update and the exit test.  The stomata call of pass 1
// sits between the two halves in the caller.
struct Loop1 {
  // per-column constants of the loop
  float SFCTMP, RHOAIR, QAIR, ZLVL, ZPD, Z0M, UR, TMPCM, TMPCH, TMPCM2, TMPCH2, Z0MG, CWP, VAIE, HCAN, rbfac, TGV, FWET,
      LAISUNE, LAISHAE, RSSUN, RSSHA, RSURF, EAIR, ESTG, GAMMAV, fveg, AIR, CIR, SAV, SFCPRS, evcap_liq, evcap_ice, THAIR, CZIL;
  // evolving
  MoninObukhov mo;
  float WSTAR, H, HG, FHG, TAH, EAH, TV;
  int LITER, iter_done, running;
  // produced by the last pass
  float RAHC, RAHG, RB, ESTV, DESTV, CAH, CVH, IRC, SHC, EVC, TR, QSFC;

  template <class Opt>
  __device__ __forceinline__ void first(const Opt& opt, int iter) {
    if (opt.sfc == 1) sfcdif1(iter, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, UR, TMPCM, TMPCH, TMPCM2, TMPCH2, mo);
    if (opt.sfc == 2) {                                                                   // func:2763-2772
      sfcdif2(iter, Z0M, TAH, THAIR, UR, CZIL, ZLVL, mo.CM, mo.CH, mo.MOZ, WSTAR, mo.FV);
      mo.CH = mo.CH / UR;
      mo.CM = mo.CM / UR;
    }
#ifdef NMP_FAST_ALGEBRA
    // The fast build (tolerance-checked, never bit-compared) merges reciprocals: RAHC is only ever used as 1/RAHC, the
    // Obukhov-length chain needs one reciprocal, x**0.5 is a square root, and /CWPC, /KH share one reciprocal.
    RAHC = fmin2(1.0f, mo.CH * UR);                                                       // holds CAH = 1/RAHC in this build
    float MOZG = 0.0f;
    if (iter > 1) {
      float TMP1 = HG * ((KARMAN * GRAV) / (TAH * (RHOAIR * CPAIR)));
      if (fabsf(TMP1) <= MPE) TMP1 = MPE;
      MOZG = fmin2((ZPD - Z0MG) * (-TMP1 / cube(mo.FV)), 1.0f);
    }
    const float FHGNEW = (MOZG < 0.0f) ? m_powf(1.0f - 15.0f * MOZG, -0.25f) : 1.0f + 4.7f * MOZG;
    FHG = (iter == 1) ? FHGNEW : 0.5f * (FHG + FHGNEW);
    const float CWPC = sqrtf(CWP * VAIE * HCAN * FHG);
    const float RHCAN = 1.0f / HCAN;
    const float TMP1 = m_expf(-CWPC * Z0MG * RHCAN);
    const float TMP2 = m_expf(-CWPC * (Z0M + ZPD) * RHCAN);
    const float KH = fmax2(KARMAN * mo.FV * (HCAN - ZPD), MPE);
    RAHG = HCAN * m_expf(CWPC) * (TMP1 - TMP2) / (CWPC * KH);
    const float TMPRB = CWPC * 50.0f / (1.0f - m_expf(-CWPC / 2.0f));
    RB = TMPRB * rbfac;
#else
    RAHC = fmax2(1.0f, 1.0f / (mo.CH * UR));                                              // func:2774-2776
    // ragrb, func:3260-3350 (Z0H = Z0M, Z0HG = Z0MG: func:2746-2752)
    float MOZG = 0.0f;
    if (iter > 1) {
      float TMP1 = KARMAN * (GRAV / TAH) * HG / (RHOAIR * CPAIR);
      if (fabsf(TMP1) <= MPE) TMP1 = MPE;
      const float MOLG = -1.0f * cube(mo.FV) / TMP1;
      MOZG = fmin2((ZPD - Z0MG) / MOLG, 1.0f);
    }
    const float FHGNEW = (MOZG < 0.0f) ? m_powf(1.0f - 15.0f * MOZG, -0.25f) : 1.0f + 4.7f * MOZG;
    FHG = (iter == 1) ? FHGNEW : 0.5f * (FHG + FHGNEW);
    const float CWPC = m_powf(CWP * VAIE * HCAN * FHG, 0.5f);
    const float TMP1 = m_expf(-CWPC * Z0MG / HCAN);
    const float TMP2 = m_expf(-CWPC * (Z0M + ZPD) / HCAN);
    const float TMPRAH2 = HCAN * m_expf(CWPC) / CWPC * (TMP1 - TMP2);
    const float KH = fmax2(KARMAN * mo.FV * (HCAN - ZPD), MPE);
    RAHG = TMPRAH2 / KH;
    const float TMPRB = CWPC * 50.0f / (1.0f - m_expf(-CWPC / 2.0f));
    RB = TMPRB * rbfac;
#endif
    const float Tc = tdc(TV);                                                             // func:2787-2795
    esat_sel(Tc, ESTV, DESTV);
  }

#ifdef NMP_FAST_ALGEBRA
  __device__ __forceinline__ void second(int iter) {
    CAH = RAHC;                                             // first() left 1/RAHC there            func:2817-2823
    const float RRB = 1.0f / RB;
    CVH = 2.0f * VAIE * RRB;
    const float CGH = 1.0f / RAHG;
    const float RCOND = 1.0f / (CAH + CVH + CGH);
    const float ATA = (SFCTMP * CAH + TGV * CGH) * RCOND;
    const float BTA = CVH * RCOND;
    const float RHOCP = RHOAIR * CPAIR;
    const float CSH = (1.0f - BTA) * RHOCP * CVH;
    const float CAW = CAH;                                                                // func:2826-2834
    const float CEW = FWET * VAIE * RRB;
    const float RSUN = RB + RSSUN, RSHA = RB + RSSHA;
    const float CTW = (1.0f - FWET) * (LAISUNE * RSHA + LAISHAE * RSUN) / (RSUN * RSHA);
    const float CGW = 1.0f / (RAHG + RSURF);
    const float RCONDW = 1.0f / (CAW + CEW + CTW + CGW);
    const float AEA = (EAIR * CAW + ESTG * CGW) * RCONDW;
    const float BEA = (CEW + CTW) * RCONDW;
    const float RCG = RHOCP / GAMMAV;
    const float CEV = (1.0f - BEA) * CEW * RCG;
    const float CTR = (1.0f - BEA) * CTW * RCG;
    TAH = ATA + BTA * TV;                                                                 // func:2837-2838
    EAH = AEA + BEA * ESTV;
    const float TV3 = cube(TV);
    IRC = fveg * (AIR + CIR * (TV3 * TV));                                                // func:2840-2848
    SHC = fveg * RHOCP * CVH * (TV - TAH);
    EVC = fveg * RCG * CEW * (ESTV - EAH);
    TR = fveg * RCG * CTW * (ESTV - EAH);
    if (TV > TFRZ) EVC = fmin2(evcap_liq, EVC);
    else EVC = fmin2(evcap_ice, EVC);
    const float B = SAV - IRC - SHC - EVC - TR;                                           // func:2850-2852
    const float A = fveg * (4.0f * CIR * TV3 + CSH + (CEV + CTR) * DESTV);
    const float DTV = B / A;
    IRC = IRC + fveg * 4.0f * CIR * TV3 * DTV;                                            // func:2854-2860
    SHC = SHC + fveg * CSH * DTV;
    EVC = EVC + fveg * CEV * DESTV * DTV;
    TR = TR + fveg * CTR * DESTV * DTV;
    TV = TV + DTV;
    H = RHOCP * (TAH - SFCTMP) * CAH;                                                     // func:2864-2868
    HG = RHOCP * (TGV - TAH) * CGH;
    iter_done = iter;                                       // QSFC (func:2868) is a function of the final EAH: see qsfc()
    if (LITER == 1) running = 0;                                                          // func:2870-2875
    else if (iter >= 5 && fabsf(DTV) <= 0.01f) LITER = 1;
  }
#else
  __device__ __forceinline__ void second(int iter) {
    CAH = 1.0f / RAHC;                                                                    // func:2817-2823
    const float RAWC = RAHC, RAWG = RAHG;
    CVH = 2.0f * VAIE / RB;
    const float CGH = 1.0f / RAHG;
    float COND = CAH + CVH + CGH;
    const float ATA = (SFCTMP * CAH + TGV * CGH) / COND;
    const float BTA = CVH / COND;
    const float CSH = (1.0f - BTA) * RHOAIR * CPAIR * CVH;
    const float CAW = 1.0f / RAWC;                                                        // func:2826-2834
    const float CEW = FWET * VAIE / RB;
    const float CTW = (1.0f - FWET) * (LAISUNE / (RB + RSSUN) + LAISHAE / (RB + RSSHA));
    const float CGW = 1.0f / (RAWG + RSURF);
    COND = CAW + CEW + CTW + CGW;
    const float AEA = (EAIR * CAW + ESTG * CGW) / COND;
    const float BEA = (CEW + CTW) / COND;
    const float CEV = (1.0f - BEA) * CEW * RHOAIR * CPAIR / GAMMAV;
    const float CTR = (1.0f - BEA) * CTW * RHOAIR * CPAIR / GAMMAV;
    TAH = ATA + BTA * TV;                                                                 // func:2837-2838
    EAH = AEA + BEA * ESTV;
    IRC = fveg * (AIR + CIR * pow4(TV));                                                  // func:2840-2848
    SHC = fveg * RHOAIR * CPAIR * CVH * (TV - TAH);
    EVC = fveg * RHOAIR * CPAIR * CEW * (ESTV - EAH) / GAMMAV;
    TR = fveg * RHOAIR * CPAIR * CTW * (ESTV - EAH) / GAMMAV;
    if (TV > TFRZ) EVC = fmin2(evcap_liq, EVC);
    else EVC = fmin2(evcap_ice, EVC);
    const float B = SAV - IRC - SHC - EVC - TR;                                           // func:2850-2852
    const float A = fveg * (4.0f * CIR * cube(TV) + CSH + (CEV + CTR) * DESTV);
    const float DTV = B / A;
    IRC = IRC + fveg * 4.0f * CIR * cube(TV) * DTV;                                       // func:2854-2860
    SHC = SHC + fveg * CSH * DTV;
    EVC = EVC + fveg * CEV * DESTV * DTV;
    TR = TR + fveg * CTR * DESTV * DTV;
    TV = TV + DTV;
    H = RHOAIR * CPAIR * (TAH - SFCTMP) / RAHC;                                           // func:2864-2868
    HG = RHOAIR * CPAIR * (TGV - TAH) / RAHG;
    QSFC = (0.622f * EAH) / (SFCPRS - 0.378f * EAH);
    iter_done = iter;
    if (LITER == 1) running = 0;                                                          // func:2870-2875
    else if (iter >= 5 && fabsf(DTV) <= 0.01f) LITER = 1;
  }
#endif
};
// Field lists for the shared-memory hand-off of a column's loop state (member-wise, so the struct itself stays in
// registers): what a pass only reads, what it carries from pass to pass, and what the caller reads after the last pass.
// Left out on purpose, so that the compiler can still drop them as dead code: mo.FM2 / mo.CH2 (the 2-m momentum
// functions, computed by func:3449-3498 and read by nobody), the first()->second() temporaries RAHC RB ESTV DESTV, and
// QSFC (a function of the final EAH, recomputed by the receiver).  THAIR CZIL WSTAR travel only for opt_sfc = 2.
#define NMP_LOOP1_CONST(X)                                                                                           \
  X(SFCTMP) X(RHOAIR) X(QAIR) X(ZLVL) X(ZPD) X(Z0M) X(UR) X(TMPCM) X(TMPCH) X(TMPCM2) X(TMPCH2) X(Z0MG) X(CWP) X(VAIE)  \
  X(HCAN) X(rbfac) X(TGV) X(FWET) X(LAISUNE) X(LAISHAE) X(RSSUN) X(RSSHA) X(RSURF) X(EAIR) X(ESTG) X(GAMMAV) X(fveg)   \
  X(AIR) X(CIR) X(SAV) X(SFCPRS) X(evcap_liq) X(evcap_ice)
#define NMP_LOOP1_SFC2(X) X(THAIR) X(CZIL) X(WSTAR)
#define NMP_LOOP1_VARF(X)                                                                                            \
  X(mo.MOZ) X(mo.FM) X(mo.FH) X(mo.FH2) X(mo.FV) X(mo.CM) X(mo.CH) X(H) X(HG) X(FHG) X(TAH) X(EAH) X(TV)
#define NMP_LOOP1_VARI(X) X(mo.MOZSGN) X(LITER) X(iter_done) X(running)
#define NMP_LOOP1_OUT(X) X(RAHG) X(CAH) X(CVH) X(IRC) X(SHC) X(EVC) X(TR)
#define NMP_COUNT(f) +1
constexpr int LOOP1_NCONST = 0 NMP_LOOP1_CONST(NMP_COUNT);
constexpr int LOOP1_NVAR = 0 NMP_LOOP1_VARF(NMP_COUNT) NMP_LOOP1_VARI(NMP_COUNT);
constexpr int LOOP1_NOUT = 0 NMP_LOOP1_OUT(NMP_COUNT);
constexpr int LOOP1_NSFC2 = 0 NMP_LOOP1_SFC2(NMP_COUNT);
static_assert(LOOP1_NOUT <= LOOP1_NCONST, "the results travel back in the rows that carried the constants");
constexpr int LOOP1_ROWS = LOOP1_NCONST + LOOP1_NVAR + LOOP1_NSFC2;

// Optional tail compaction (-DNMP_COMPACT_LOOP1): after NMP_COMPACT_AFTER passes the columns that still iterate are
// packed into the first warps of the CTA through shared memory, so the rest of the tail runs on full warps.  Up to
// NMP_COMPACT_CAP columns per CTA travel; the rest keep iterating in place.
#ifdef NMP_COMPACT_LOOP1
#ifndef NMP_COMPACT_CAP
#define NMP_COMPACT_CAP 128
#endif
#ifndef NMP_COMPACT_AFTER
#define NMP_COMPACT_AFTER 8
#endif
constexpr int LOOP1_CAP = NMP_COMPACT_CAP;
// One packed column: its own function, so the register allocation of the kernel body does not see a second Loop1.
template <class Opt>
__device__ __noinline__ void loop1_job(const Opt& opt, float (*xch)[LOOP1_CAP], int t) {
  Loop1 W;
  W.mo.FM2 = 0.0f; W.mo.CH2 = 0.0f; W.THAIR = 0.0f; W.CZIL = 0.0f; W.WSTAR = 0.0f;   // dead (or opt_sfc = 2 only)
  int w = 0;
#define X(fld) W.fld = xch[w++][t];
  NMP_LOOP1_CONST(X) NMP_LOOP1_VARF(X)
#undef X
#define X(fld) W.fld = __float_as_int(xch[w++][t]);
  NMP_LOOP1_VARI(X)
#undef X
  if (opt.sfc == 2) {
#define X(fld) W.fld = xch[w++][t];
    NMP_LOOP1_SFC2(X)
#undef X
  }
#pragma unroll 1
  for (int iter = NMP_COMPACT_AFTER + 1; iter <= 20; ++iter) {
    if (!W.running) break;
    W.first(opt, iter);
    W.second(iter);
  }
  w = 0;
#define X(fld) xch[w++][t] = W.fld;
  NMP_LOOP1_VARF(X)
#undef X
#define X(fld) xch[w++][t] = __int_as_float(W.fld);
  NMP_LOOP1_VARI(X)
#undef X
#define X(fld) xch[w++][t] = W.fld;
  NMP_LOOP1_OUT(X)
#undef X
  if (opt.sfc == 2) xch[LOOP1_ROWS - 1][t] = W.WSTAR;
}
#endif

// ---- func:322-328  layer thickness from the interface depths ------------------------------------------
__device__ __forceinline__ void layer_thickness(const Column& c, float (&DZSNSO)[7]) {
  static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
    if (k == c.ISNOW + 1) DZSNSO[IX7(k)] = -c.ZSNSO[IX7(k)];
    else if (k > c.ISNOW + 1) DZSNSO[IX7(k)] = c.ZSNSO[IX7(k > -2 ? k - 1 : k)] - c.ZSNSO[IX7(k)];
    else DZSNSO[IX7(k)] = 0.0f;
  });
}
// The same seven values again, after a stretch of code that does not use them (the flux solvers): the empty asm makes
// the depths opaque, so the optimiser cannot merge this with the first evaluation and keep seven registers (or
// spill slots) alive across the Newton loops.  No instruction is generated for it.
__device__ __forceinline__ void layer_thickness_again(Column& c, float (&DZSNSO)[7]) {
  static_for<0, 6>([&](auto k_) { constexpr int k = decltype(k_)::value; asm volatile("" : "+f"(c.ZSNSO[k])); });
  layer_thickness(c, DZSNSO);
}

// =====================================================================================================
// energy: func:735-1338 with its callees thermoprop/csnow (1341-1497), radiation/albedo/surrad/snowage/
// snowalb_*/groundalb (1598-2212), vege_flux (2465-2964), bare_flux (2967-3257), ragrb (3260-3350),
// tsnosoi/hrt/hstep/rosr12 (3987-4288) and phasechange/frh2o (4291-4598).
// =====================================================================================================
template <class Opt, bool DIAG>
__device__ __forceinline__ void energy(const Opt& opt, const Tables& T, const RunConst& rc, Column& c,
                                       const Forcing& f, float RHOAIR, float QAIR, float THAIR, float EAIR,
                                       float SWDOWN, float IGS, float HTOP, float elai, float esai, float fveg,
                                       const float (&DZSNSO_in)[7], EnergyResult& r, const DiagSink<DIAG>& dg) {
  const float* VG = &T.veg[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGF];
  const float* SP = &T.soil[NMP_CHK(c.sltyp - 1, 0, MAXSOILTYP - 1, 4) * NSOILP];
  const int ISURBAN = T.flags[0];
  const float DT = rc.dt;
  const int ISNOW = c.ISNOW;
  float DZSNSO[7];   // re-derived before tsnosoi (layer_thickness_again), not kept across the flux solvers
  static_for<0, 6>([&](auto k_) { constexpr int k = decltype(k_)::value; DZSNSO[k] = DZSNSO_in[k]; });
  const float SMCMAX = SP[P_SMCMAX], BEXP = SP[P_BEXP], PSISAT = SP[P_PSISAT], SMCWLT = SP[P_SMCWLT];
  const float Z0 = 0.01f, PSIWLT = -150.0f;                                               // func:1017-1018
  r.flags = 0;
  r.iter_vege = 0; r.iter_stom_sun = 0; r.iter_stom_sha = 0;

  const float UR = fmax2(sqrtf(f.UU * f.UU + f.VV * f.VV), 1.0f);                         // func:1041
  const float VAI = elai + esai;                                                          // func:1044-1046
  const bool VEG = VAI > 0.0f;

  float FSNO = 0.0f;                                                                      // func:1049-1054
  if (c.snowh > 0.0f) {
    const float BDSNO = c.sneqv / c.snowh;
    const float FMELT = m_powf(BDSNO / 100.0f, T.gen[G_MLTFCT]);
    FSNO = m_tanhf(c.snowh / (2.5f * Z0 * FMELT));
  }
  float Z0MG;                                                                             // func:1057-1065
  if (c.IST == 2) Z0MG = (c.TG <= TFRZ) ? 0.01f * (1.0f - FSNO) + FSNO * T.gen[G_Z0SNO] : 0.01f;
  else Z0MG = Z0 * (1.0f - FSNO) + FSNO * T.gen[G_Z0SNO];
  const float ZPDG = c.snowh;                                                             // func:1069-1077
  float Z0M, ZPD;
  if (VEG) {
    Z0M = VG[V_Z0MVT];
    ZPD = 0.65f * HTOP;
    if (c.snowh > ZPD) ZPD = c.snowh;
  } else {
    Z0M = Z0MG;
    ZPD = ZPDG;
  }
  float ZLVL = fmax2(ZPD, HTOP) + rc.zref;                                                // func:1079-1080
  if (ZPDG >= ZLVL) ZLVL = ZPDG + rc.zref;
  const float CWP = VG[V_CWPVT];                                                          // func:1084

  NMP_PHASE_SYNC_AT(3);
  // ---- thermoprop, func:1341-1445 (+ csnow 1448-1497, tdfcnd) -----------------------------------
  // Only the top layer's conductivity is needed by the flux solvers (func:1215/1231): the snow layers' and soil layer 1's
  // (with its interface blend, func:1440-1444) are evaluated here, layers 2-4, the heat capacities and FACT just before
  // tsnosoi - functions of layer state that nothing touches in between (same values, fewer words alive across the loops).
  float DF1, DFtop;
  {
    float DFS[3];
    static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
      DFS[IX3(k)] = 0.0f;
      if (k >= ISNOW + 1) {
        const float BDSNOI = (c.SNICE[IX3(k)] + c.SNLIQ[IX3(k)]) / DZSNSO[IX7(k)];        // func:1482-1490
        DFS[IX3(k)] = 3.2217E-6f * (BDSNOI * BDSNOI);
      }
    });
    NMP_PHASE_SYNC_AT(4);
    DF1 = tdfcnd(SMCMAX, SP[P_QUARTZ], c.SMC[0], c.soilwat[0]);                           // func:1398-1403
    if (c.lutyp == ISURBAN) DF1 = 3.24f;                                                  // func:1405-1409
    if (c.IST == 2) DF1 = (c.STC[IX7(1)] > TFRZ) ? TKWAT : TKICE;                         // func:1420-1430
    if (ISNOW == 0) DF1 = (DF1 * DZSNSO[IX7(1)] + 0.35f * c.snowh) / (c.snowh + DZSNSO[IX7(1)]);  // func:1440-1444
    else DF1 = (DF1 * DZSNSO[IX7(1)] + DFS[IX3(0)] * DZSNSO[IX7(0)]) / (DZSNSO[IX7(0)] + DZSNSO[IX7(1)]);
    DFtop = (ISNOW == 0) ? DF1 : ((ISNOW == -1) ? DFS[IX3(0)] : ((ISNOW == -2) ? DFS[IX3(-1)] : DFS[IX3(-2)]));
  }

  NMP_PHASE_SYNC_AT(5);
  // ---- radiation, func:1598-1714: albedo (1717-1887) + surrad (1890-2005) ----------------------
  float FSUN = 0.0f, BGAP = 0.0f, WGAP = 0.0f;
  float ALBGRD[2] = {0.0f, 0.0f}, ALBGRI[2] = {0.0f, 0.0f}, ALBD[2] = {0.0f, 0.0f}, ALBI[2] = {0.0f, 0.0f};
  float FABD[2] = {0.0f, 0.0f}, FABI[2] = {0.0f, 0.0f}, FTDD[2] = {0.0f, 0.0f}, FTID[2] = {0.0f, 0.0f};
  float FTII[2] = {0.0f, 0.0f};
  float FREVD[2] = {0.0f, 0.0f}, FREVI[2] = {0.0f, 0.0f}, FREGD[2] = {0.0f, 0.0f}, FREGI[2] = {0.0f, 0.0f};
  const bool day = f.COSZ > 0.0f;                                                         // func:1823 (night: all zero)
  float RHO[2] = {0.0f, 0.0f}, TAU[2] = {0.0f, 0.0f};
  TwoStreamCommon ts;
  ts.GDIR = 0.0f;
  if (day) {
    const float WL = elai / fmax2(VAI, MPE), WS = esai / fmax2(VAI, MPE);                 // func:1827-1833
    static_for<0, (2) - 1>([&](auto ib_) { constexpr int ib = decltype(ib_)::value;
      RHO[ib] = fmax2(VG[V_RHOL + ib] * WL + VG[V_RHOS + ib] * WS, MPE);
      TAU[ib] = fmax2(VG[V_TAUL + ib] * WL + VG[V_TAUS + ib] * WS, MPE);
    });
    // snowage, func:2008-2054
    float FAGE;
    if (c.sneqv <= 0.0f || c.sneqv > 800.0f) {
      c.TAUSS = 0.0f;
    } else {
      const float DELA0 = 1.0E-6f * DT;
      const float ARG = 5.0E3f * (1.0f / TFRZ - 1.0f / c.TG);
      const float TAGE = m_expf(ARG) + m_expf(fmin2(0.0f, 10.0f * ARG)) + 0.3f;
      const float DELA = DELA0 * TAGE;
      const float DELS = fmax2(0.0f, c.sneqv - c.SNEQVO) / T.gen[G_SWEMAX];
      c.TAUSS = fmax2(0.0f, (c.TAUSS + DELA) * (1.0f - DELS));
    }
    FAGE = c.TAUSS / (c.TAUSS + 1.0f);
    float ALBSND[2] = {0.0f, 0.0f}, ALBSNI[2] = {0.0f, 0.0f};
    if (opt.alb == 1) {                                                                   // snowalb_bats, func:2057-2102
      const float SL1 = 1.0f / 2.0f, SL2 = 2.0f * 2.0f;
      const float FZEN = fmax2(((1.0f + SL1) / (1.0f + SL2 * f.COSZ) - SL1), 0.0f);
      ALBSNI[0] = 0.95f * (1.0f - 0.2f * FAGE);
      ALBSNI[1] = 0.65f * (1.0f - 0.5f * FAGE);
      ALBSND[0] = ALBSNI[0] + 0.4f * FZEN * (1.0f - ALBSNI[0]);
      ALBSND[1] = ALBSNI[1] + 0.4f * FZEN * (1.0f - ALBSNI[1]);
    }
    if (opt.alb == 2) {                                                                   // snowalb_class, func:2105-2151
      float ALB = 0.55f + (c.ALBOLD - 0.55f) * m_expf(-0.01f * DT / 3600.0f);
      if (c.QSNOW > 0.0f) ALB = ALB + fmin2(c.QSNOW * DT, T.gen[G_SWEMAX]) * (0.84f - ALB) / T.gen[G_SWEMAX];
      ALBSNI[0] = ALB; ALBSNI[1] = ALB; ALBSND[0] = ALB; ALBSND[1] = ALB;
      c.ALBOLD = ALB;                                                                     // func:1845
    }
    // groundalb, func:2154-2212
    const float INC = fmax2(0.11f - 0.40f * c.SMC[0], 0.0f);
    static_for<0, (2) - 1>([&](auto ib_) { constexpr int ib = decltype(ib_)::value;
      float ALBSOD, ALBSOI;
      if (c.IST == 1) {
        ALBSOD = fmin2(T.color[NMP_CHK(c.ISC - 1, 0, MAXCOLOR - 1, 4) * NCOLORP + K_ALBSAT + ib] + INC, T.color[NMP_CHK(c.ISC - 1, 0, MAXCOLOR - 1, 4) * NCOLORP + K_ALBDRY + ib]);
        ALBSOI = ALBSOD;
      } else if (c.TG > TFRZ) {
        ALBSOD = 0.06f / (m_powf(fmax2(0.01f, f.COSZ), 1.7f) + 0.15f);
        ALBSOI = 0.06f;
      } else {
        ALBSOD = T.gen[G_ALBLAKE + ib];
        ALBSOI = ALBSOD;
      }
      if (c.IST == 1 && c.ISC == 9) { ALBSOD = ALBSOD + 0.10f; ALBSOI = ALBSOI + 0.10f; }
      ALBGRD[ib] = ALBSOD * (1.0f - FSNO) + ALBSND[ib] * FSNO;
      ALBGRI[ib] = ALBSOI * (1.0f - FSNO) + ALBSNI[ib] * FSNO;
    });
    // twostream x (2 bands) x (direct, diffuse), func:1858-1873
    twostream_common(opt.rad, f.COSZ, VAI, fveg, VG[V_XL], VG[V_RCROWN], VG[V_HVT], VG[V_HVB], BGAP, WGAP, ts);
    BGAP = ts.BGAP;
    WGAP = ts.WGAP;
  }
#pragma unroll 1
  for (int call = 0; call < 4; ++call) {
    NMP_PHASE_SYNC_AT(6);
    if (day) {
      const int ib = call >> 1, ic = call & 1;
      TwoStreamOut o;
      twostream_band(ts, ic, VAI, c.FWET, c.TV, ib ? ALBGRD[1] : ALBGRD[0], ib ? ALBGRI[1] : ALBGRI[0],
                     ib ? RHO[1] : RHO[0], ib ? TAU[1] : TAU[0], T.gen[G_OMEGAS + ib], T.gen[G_BETADS],
                     T.gen[G_BETAIS], o);
      static_for<0, (2) - 1>([&](auto b_) { constexpr int b = decltype(b_)::value;
        if (ib == b && ic == 0) { FABD[b] = o.FAB; ALBD[b] = o.FRE; FTDD[b] = o.FTD; FTID[b] = o.FTI; FREVD[b] = o.FREV; FREGD[b] = o.FREG; }
        if (ib == b && ic == 1) { FABI[b] = o.FAB; ALBI[b] = o.FRE; FTII[b] = o.FTI; FREVI[b] = o.FREV; FREGI[b] = o.FREG; }
      });
    }
  }
  if (day) {
    // sunlit fraction, func:1877-1886
    const float EXT = ts.GDIR / f.COSZ * sqrtf(1.0f - RHO[0] - TAU[0]);
    FSUN = (1.0f - m_expf(-EXT * VAI)) / fmax2(EXT * VAI, MPE);
    if (FSUN < 0.01f) FSUN = 0.0f;
  }
  const float FSHA = 1.0f - FSUN;                                                         // func:1695-1698
  const float LAISUN = elai * FSUN, LAISHA = elai * FSHA;
  // surrad, func:1890-2005
  float SAG = 0.0f, SAV = 0.0f, FSA = 0.0f, PARSUN, PARSHA, FSR, FSRV, FSRG;
  const float SOLAD[2] = {SWDOWN * 0.7f * 0.5f, SWDOWN * 0.7f * 0.5f};                    // func:526-529
  const float SOLAI[2] = {SWDOWN * 0.3f * 0.5f, SWDOWN * 0.3f * 0.5f};
  {
    float CAD[2], CAI[2];
    static_for<0, (2) - 1>([&](auto ib_) { constexpr int ib = decltype(ib_)::value;
      CAD[ib] = SOLAD[ib] * FABD[ib];
      CAI[ib] = SOLAI[ib] * FABI[ib];
      SAV = SAV + CAD[ib] + CAI[ib];
      FSA = FSA + CAD[ib] + CAI[ib];
      const float TRD = SOLAD[ib] * FTDD[ib];
      const float TRI = SOLAD[ib] * FTID[ib] + SOLAI[ib] * FTII[ib];
      const float ABS = TRD * (1.0f - ALBGRD[ib]) + TRI * (1.0f - ALBGRI[ib]);
      SAG = SAG + ABS;
      FSA = FSA + ABS;
    });
    const float LAIFRA = elai / fmax2(VAI, MPE);
    if (FSUN > 0.0f) {
      PARSUN = (CAD[0] + FSUN * CAI[0]) * LAIFRA / fmax2(LAISUN, MPE);
      PARSHA = (FSHA * CAI[0]) * LAIFRA / fmax2(LAISHA, MPE);
    } else {
      PARSUN = 0.0f;
      PARSHA = (CAD[0] + CAI[0]) * LAIFRA / fmax2(LAISHA, MPE);
    }
    const float RVIS = ALBD[0] * SOLAD[0] + ALBI[0] * SOLAI[0];
    const float RNIR = ALBD[1] * SOLAD[1] + ALBI[1] * SOLAI[1];
    FSR = RVIS + RNIR;
    FSRV = FREVD[0] * SOLAD[0] + FREVI[0] * SOLAI[0] + FREVD[1] * SOLAD[1] + FREVI[1] * SOLAI[1];
    FSRG = FREGD[0] * SOLAD[0] + FREGI[0] * SOLAI[0] + FREGD[1] * SOLAD[1] + FREGI[1] * SOLAI[1];
  }
  r.SAV = SAV; r.SAG = SAG; r.FSA = FSA; r.FSR = FSR; r.FSRV = FSRV; r.FSRG = FSRG;
  dg.put(D_BGAP, BGAP);
  dg.put(D_WGAP, WGAP);
  dg.put(D_FSNO, FSNO);

  NMP_PHASE_SYNC_AT(7);
  // ---- emissivities, soil-moisture stress, surface resistance: func:1106-1189 -------------------
  const float EMV = 1.0f - m_expf(-(elai + esai) / 1.0f);
  float EMG;
  if (c.ICE == 1) EMG = 0.98f * (1.0f - FSNO) + 1.0f * FSNO;
  else if (c.IST == 1) EMG = T.gen[G_EMSSOIL] * (1.0f - FSNO) + 1.0f * FSNO;
  else EMG = T.gen[G_EMSLAKE] * (1.0f - FSNO) + 1.0f * FSNO;

  const int NROOT = T.vegi[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGI + VI_NROOT];
  float BTRAN = 0.0f;
  static_for<0, (4) - 1>([&](auto k_) { constexpr int k = decltype(k_)::value;
    r.BTRANI[k] = 0.0f;
  });
  if (c.IST == 1) {                                                                       // func:1119-1140
    const float zroot = root_depth(rc, NROOT);
    static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      if (k <= NROOT) {
        float GX = 0.0f;
        if (opt.btr == 1) GX = (c.soilwat[IX4(k)] - SMCWLT) / (SP[P_SMCREF] - (SMCWLT));
        if (opt.btr == 2 || opt.btr == 3) {
          const float PSI = fmax2(PSIWLT, -PSISAT * m_powf(fmax2(0.01f, c.soilwat[IX4(k)]) / SMCMAX, -BEXP));
          if (opt.btr == 2) GX = (1.0f - PSI / PSIWLT) / (1.0f + PSISAT / PSIWLT);
          else GX = 1.0f - m_expf(-5.8f * (m_logf(PSIWLT / PSI)));
        }
        GX = fmin2(1.0f, fmax2(0.0f, GX));
        r.BTRANI[IX4(k)] = fmax2(MPE, DZSNSO[IX7(k)] / zroot * GX);
        BTRAN = BTRAN + r.BTRANI[IX4(k)];
      }
    });
    BTRAN = fmax2(MPE, BTRAN);
    static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      if (k <= NROOT) r.BTRANI[IX4(k)] = r.BTRANI[IX4(k)] / BTRAN;
    });
  }
  r.BTRAN = BTRAN;

  float RSURF, RHSUR;
  if (c.IST == 2) {                                                                       // func:1144-1164
    RSURF = 1.0f;
    RHSUR = 1.0f;
  } else {
    const float L_RSURF = (-rc.zsoil[0]) * (m_expf(pow5(1.0f - fmin2(1.0f, c.soilwat[0] / SMCMAX))) - 1.0f) /
                          (2.71828f - 1.0f);
    const float D_RSURF = 2.2E-5f * SMCMAX * SMCMAX * m_powf(1.0f - SMCWLT / SMCMAX, 2.0f + 3.0f / BEXP);
    RSURF = L_RSURF / D_RSURF;
    if (c.soilwat[0] < 0.01f && c.snowh == 0.0f) RSURF = 1.0E6f;
    const float PSI = -PSISAT * m_powf(fmax2(0.01f, c.soilwat[0]) / SMCMAX, -BEXP);
    RHSUR = FSNO + (1.0f - FSNO) * m_expf(PSI * GRAV / (RVAP * c.TG));
  }
  if (c.lutyp == ISURBAN && c.snowh == 0.0f) RSURF = 1.0E6f;                              // func:1167-1169

  const float LATHEAV = (c.TV > TFRZ) ? HVAP : HSUB;                                      // func:1173-1189
  r.frozen_canopy = !(c.TV > TFRZ);
  const float GAMMAV = CPAIR * f.SFCPRS / (0.622f * LATHEAV);
  const float LATHEAG = (c.TG > TFRZ) ? HVAP : HSUB;
  r.frozen_ground = !(c.TG > TFRZ);
  const float GAMMAG = CPAIR * f.SFCPRS / (0.622f * LATHEAG);
  r.LATHEAG = LATHEAG;

  const float PSFC = f.SFCPRS;   // offline driver: lowest-model-level pressure == surface pressure
  const float STCtop = pick(c.STC, IX7(ISNOW + 1)), DZtop = pick(DZSNSO, IX7(ISNOW + 1));
  const float CO2AIR = rc.co2_frac * f.SFCPRS, O2AIR = rc.o2_frac * f.SFCPRS;

  // outputs of the two tiles (zeros of func:1021-1038 for the canopy tile)
  float TAUXV = 0.0f, TAUYV = 0.0f, IRC = 0.0f, SHC = 0.0f, IRG = 0.0f, SHG = 0.0f, EVG = 0.0f, EVC = 0.0f, TR = 0.0f;
  float GHV = 0.0f, PSNSUN = 0.0f, PSNSHA = 0.0f, T2MV = 0.0f, Q2V = 0.0f, CHV = 0.0f, CHLEAF = 0.0f, CHUC = 0.0f;
  float CHV2 = 0.0f, TGV = 0.0f, CMV = 0.0f, RSSUN = 0.0f, RSSHA = 0.0f;
  float QSFC = c.QSFC;

  NMP_PHASE_SYNC_AT(8);
  // ---- vege_flux, func:2465-2964 ----------------------------------------------------------------
  // The three Newton loops run as CTA-uniform loops (NMP_ANY = __syncthreads_or when phase barriers are
  // on): every thread executes the loop control, the body is predicated on the lane still iterating.  Per
  // column the sequence of operations is exactly the reference's.
  const bool veg_tile = VEG && fveg > 0.0f;                                               // func:1200
  float TAH = c.TAH, EAH = c.EAH, TV = c.TV;
  MoninObukhov mo;
  mo.MOZ = 0.0f; mo.FM = 0.0f; mo.FH = 0.0f; mo.FM2 = 0.0f; mo.FH2 = 0.0f; mo.FV = 0.1f; mo.CM = c.CM; mo.CH = c.CH;
  mo.CH2 = 0.0f; mo.MOZSGN = 0;
  float RAHG = 0.0f, RAWG = 0.0f;
  float ESTG = 0.0f, CAH = 0.0f, CVH = 0.0f;
  float HCAN = HTOP, AIR = 0.0f, CIR = 0.0f;
  float CSH = 0.0f, CEV = 0.0f, CGH = 0.0f;
  const float Z0H = Z0M;                                                                  // func:2746-2752
  LeafParams LP;
  LP.BP = VG[V_BP]; LP.MP = VG[V_MP]; LP.FOLNMX = VG[V_FOLNMX]; LP.QE25 = VG[V_QE25]; LP.KC25 = VG[V_KC25];
  LP.AKC = VG[V_AKC]; LP.KO25 = VG[V_KO25]; LP.AKO = VG[V_AKO]; LP.VCMX25 = VG[V_VCMX25]; LP.AVCMX = VG[V_AVCMX];
  LP.C3C4 = T.vegi[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGI + VI_C3C4];
  Loop1 L;
  L.mo = mo; L.WSTAR = 0.0f; L.H = 0.0f; L.HG = 0.0f; L.FHG = 0.0f; L.TAH = TAH; L.EAH = EAH; L.TV = TV;
  L.LITER = 0; L.iter_done = 0; L.running = veg_tile ? 1 : 0;
  L.RAHC = 1.0f; L.RAHG = 0.0f; L.RB = 0.0f; L.ESTV = 0.0f; L.DESTV = 0.0f; L.CAH = 0.0f; L.CVH = 0.0f; L.IRC = 0.0f;
  L.SHC = 0.0f; L.EVC = 0.0f; L.TR = 0.0f; L.QSFC = QSFC;
  L.SFCTMP = f.SFCTMP; L.RHOAIR = RHOAIR; L.QAIR = QAIR; L.ZLVL = ZLVL; L.ZPD = ZPD; L.Z0M = Z0M; L.UR = UR;
  L.TMPCM = 0.0f; L.TMPCH = 0.0f; L.TMPCM2 = 0.0f; L.TMPCH2 = 0.0f; L.Z0MG = Z0MG; L.CWP = CWP; L.VAIE = 0.0f; L.HCAN = HCAN;
  L.rbfac = 0.0f; L.TGV = c.TG; L.FWET = c.FWET; L.LAISUNE = 0.0f; L.LAISHAE = 0.0f; L.RSSUN = 0.0f; L.RSSHA = 0.0f;
  L.RSURF = RSURF; L.EAIR = EAIR; L.ESTG = 0.0f; L.GAMMAV = GAMMAV; L.fveg = fveg; L.AIR = 0.0f; L.CIR = 0.0f; L.SAV = SAV;
  L.SFCPRS = f.SFCPRS; L.evcap_liq = c.CANLIQ * LATHEAV / DT; L.evcap_ice = c.CANICE * LATHEAV / DT; L.THAIR = THAIR;
  L.CZIL = T.gen[G_CZIL];
  if (veg_tile) {
    TGV = c.TG;
    CMV = c.CM;
    CHV = c.CH;
    L.VAIE = fmin2(6.0f, VAI / fveg);                                                     // func:2704-2706
    L.LAISUNE = fmin2(6.0f, LAISUN / fveg);
    L.LAISHAE = fmin2(6.0f, LAISHA / fveg);
    const float Tc = tdc(TGV);                                                            // func:2710-2716
    { float unused; esat_sel(Tc, ESTG, unused); }
    L.ESTG = ESTG;
    L.QSFC = 0.622f * EAIR / (PSFC - 0.378f * EAIR);                                      // func:2720
    const float UC = UR * m_logf(HCAN / Z0M) / m_logf(ZLVL / Z0M);                        // func:2725
    if ((HCAN - ZPD) <= 0.0f) r.flags |= FLAG_HCAN;                                       // func:2726-2738
    if (ZLVL <= ZPD) r.flags |= FLAG_ZLVL;                                                // func:3412-3415
    L.AIR = -EMV * (1.0f + (1.0f - EMV) * (1.0f - EMG)) * f.LWDN - EMV * EMG * SB * pow4(TGV);  // func:2741-2742
    L.CIR = (2.0f - EMV * (1.0f - EMG)) * EMV * SB;
    L.TMPCM = m_logf((ZLVL - ZPD) / Z0M); L.TMPCH = m_logf((ZLVL - ZPD) / Z0H);           // func:3417-3420
    L.TMPCM2 = m_logf((2.0f + Z0M) / Z0M); L.TMPCH2 = m_logf((2.0f + Z0H) / Z0H);
    L.rbfac = sqrtf(VG[V_DLEAF] / UC);                                                    // func:3347 (loop invariant)
  }
  // ---- loop1, func:2744-2877.  The exit test cannot fire before the end of pass 6 (LITER is set at
  // pass >= 5 and acts one pass later), so passes 1..6 run for every canopy column; the data-dependent tail
  // (passes 7..20; on the CONUS workload half of the column-steps need at least one, 0.4% all fourteen) is what
  // makes warps diverge: 22.8% of the kernel's warp-instructions at 12.3 lanes active (profiles/r01_e).
#pragma unroll 1
  for (int iter = 1; iter <= 6; ++iter) {
    if (L.running) L.first(opt, iter);
    if (iter == 1) {                                                                      // func:2798-2814
      if (opt.crs == 1) {
#pragma unroll 1
        for (int leaf = 0; leaf < 2; ++leaf) {
          Stomata sm;
          sm.done = true; sm.niter = 0; sm.rs = 0.0f; sm.psn = 0.0f;
          if (L.running) sm.begin(LP, f.SFCPRS, f.SFCTMP, leaf ? PARSHA : PARSUN, L.TV, O2AIR, CO2AIR, rc.foln, BTRAN, L.RB);
#pragma unroll 1
          for (int bis = 0; bis < 20; ++bis) {
            if (!(L.running && !sm.done)) break;
            sm.step(LP, IGS, f.SFCPRS, L.EAH, L.ESTV, CO2AIR);
          }
          if (L.running) {
            if (leaf == 0) { RSSUN = sm.rs; PSNSUN = sm.psn; r.iter_stom_sun = sm.niter; }
            else { RSSHA = sm.rs; PSNSHA = sm.psn; r.iter_stom_sha = sm.niter; }
          }
        }
      }
      if (opt.crs == 2) {
        if (L.running) {
          canres(VG[V_RGL], VG[V_RSMIN], VG[V_RSMAX], VG[V_TOPT], VG[V_HS], f.SFCPRS, L.TV, PARSUN, L.EAH, BTRAN, RSSUN, PSNSUN);
          canres(VG[V_RGL], VG[V_RSMIN], VG[V_RSMAX], VG[V_TOPT], VG[V_HS], f.SFCPRS, L.TV, PARSHA, L.EAH, BTRAN, RSSHA, PSNSHA);
        }
      }
      L.RSSUN = RSSUN;
      L.RSSHA = RSSHA;
    }
    if (L.running) L.second(iter);
  }
  // ---- loop1 tail, passes 7..20 (only the lanes still iterating) --------------------------------------
#ifdef NMP_COMPACT_LOOP1
#pragma unroll 1
  for (int iter = 7; iter <= NMP_COMPACT_AFTER; ++iter) {
    if (!L.running) break;
    L.first(opt, iter);
    L.second(iter);
  }
  bool over_capacity = false;
  {
    __shared__ float xch[LOOP1_ROWS][LOOP1_CAP];
    __shared__ int xcnt;
    if (threadIdx.x == 0) xcnt = 0;
    __syncthreads();
    const unsigned ball = __ballot_sync(0xffffffffu, L.running != 0);
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0 && ball) base = atomicAdd(&xcnt, __popc(ball));
    base = __shfl_sync(0xffffffffu, base, 0);
    const int slot = base + __popc(ball & ((1u << lane) - 1u));
    const bool moved = (L.running != 0) && slot < LOOP1_CAP;
    over_capacity = (L.running != 0) && !moved;
    if (moved) {
      int w = 0;
#define X(fld) xch[w++][slot] = L.fld;
      NMP_LOOP1_CONST(X) NMP_LOOP1_VARF(X)
#undef X
#define X(fld) xch[w++][slot] = __int_as_float(L.fld);
      NMP_LOOP1_VARI(X)
#undef X
      if (opt.sfc == 2) {
#define X(fld) xch[w++][slot] = L.fld;
        NMP_LOOP1_SFC2(X)
#undef X
      }
    }
    __syncthreads();
    const int njobs = xcnt < LOOP1_CAP ? xcnt : LOOP1_CAP;
    if ((int)threadIdx.x < njobs) loop1_job(opt, xch, (int)threadIdx.x);
    __syncthreads();
    if (moved) {
      int w = 0;
#define X(fld) L.fld = xch[w++][slot];
      NMP_LOOP1_VARF(X)
#undef X
#define X(fld) L.fld = __float_as_int(xch[w++][slot]);
      NMP_LOOP1_VARI(X)
#undef X
#define X(fld) L.fld = xch[w++][slot];
      NMP_LOOP1_OUT(X)
#undef X
      if (opt.sfc == 2) L.WSTAR = xch[LOOP1_ROWS - 1][slot];
      L.QSFC = (0.622f * L.EAH) / (L.SFCPRS - 0.378f * L.EAH);                            // func:2868, as in second()
    }
  }
  // over capacity (rare): the column keeps iterating in place
#pragma unroll 1
  for (int iter = NMP_COMPACT_AFTER + 1; iter <= 20; ++iter) {
    if (!(over_capacity && L.running)) break;
    L.first(opt, iter);
    L.second(iter);
  }
#else
#ifndef NMP_EXPERIMENT_NITERC   // timing-only experiments cap the pass count (results wrong): what is the tail worth?
#define NMP_EXPERIMENT_NITERC 20
#endif
#pragma unroll 1
  for (int iter = 7; iter <= NMP_EXPERIMENT_NITERC; ++iter) {
    if (!L.running) break;
    L.first(opt, iter);
    L.second(iter);
  }
#endif
  if (veg_tile) {
#ifdef NMP_FAST_ALGEBRA
    L.QSFC = (0.622f * L.EAH) / (L.SFCPRS - 0.378f * L.EAH);                              // func:2868, of the last pass
#endif
    r.iter_vege = L.iter_done;
    mo = L.mo; TAH = L.TAH; EAH = L.EAH; TV = L.TV; RAHG = L.RAHG; RAWG = L.RAHG; CAH = L.CAH; CVH = L.CVH;
    IRC = L.IRC; SHC = L.SHC; EVC = L.EVC; TR = L.TR; QSFC = L.QSFC; AIR = L.AIR; CIR = L.CIR;
  }
  NMP_PHASE_SYNC_AT(9);
  // under-canopy ground temperature, loop2, func:2881-2914
  if (veg_tile) {
    AIR = -EMG * (1.0f - EMV) * f.LWDN - EMG * EMV * SB * pow4(TV);
    CIR = EMG * SB;
    CSH = RHOAIR * CPAIR / RAHG;
    CEV = RHOAIR * CPAIR / (GAMMAG * (RAWG + RSURF));
    CGH = 2.0f * DFtop / DZtop;
#pragma unroll 1
    for (int iter = 1; iter <= 5; ++iter) {
      const float Tc = tdc(TGV);
      float DESTG;
      esat_sel(Tc, ESTG, DESTG);
      IRG = CIR * pow4(TGV) + AIR;
      SHG = CSH * (TGV - TAH);
      EVG = CEV * (ESTG * RHSUR - EAH);
      GHV = CGH * (TGV - STCtop);
      const float B = SAG - IRG - SHG - EVG - GHV;
      const float A = 4.0f * CIR * cube(TGV) + CSH + CEV * DESTG + CGH;
      const float DTG = B / A;
      IRG = IRG + 4.0f * CIR * cube(TGV) * DTG;
      SHG = SHG + CSH * DTG;
      EVG = EVG + CEV * DESTG * DTG;
      GHV = GHV + CGH * DTG;
      TGV = TGV + DTG;
    }
    if (opt.stc == 1) {                                                                   // func:2920-2928
      if (c.snowh > 0.05f && TGV > TFRZ) {
        TGV = TFRZ;
        IRG = CIR * pow4(TGV) - EMG * (1.0f - EMV) * f.LWDN - EMG * EMV * SB * pow4(TV);
        SHG = CSH * (TGV - TAH);
        EVG = CEV * (ESTG * RHSUR - EAH);
        GHV = SAG - (IRG + SHG + EVG);
      }
    }
    TAUXV = -RHOAIR * mo.CM * UR * f.UU;                                                  // func:2932-2933
    TAUYV = -RHOAIR * mo.CM * UR * f.VV;
    {                                                                                     // func:2943-2957
      CHV2 = mo.FV * KARMAN / (m_logf((2.0f + Z0H) / Z0H) - mo.FH2);
      if (CHV2 < 1.E-5f) {
        T2MV = TAH;
        Q2V = QSFC;
      } else {
        T2MV = TAH - (SHG + SHC / fveg) / (RHOAIR * CPAIR) * 1.0f / CHV2;
        Q2V = QSFC - ((EVC + TR) / fveg + EVG) / (LATHEAV * RHOAIR) * 1.0f / CHV2;
      }
    }
    CMV = mo.CM;
    CHV = CAH;                                                                            // func:2960-2962
    CHLEAF = CVH;
    CHUC = 1.0f / RAHG;
    c.TAH = TAH;
    c.EAH = EAH;
    c.TV = TV;
  }

  NMP_PHASE_SYNC_AT(10);
  // ---- bare_flux, func:2967-3257 ----------------------------------------------------------------
  float TGB = c.TG, CMB = c.CM, CHB = c.CH, TAUXB, TAUYB, IRB = 0.0f, SHB = 0.0f, EVB = 0.0f, GHB = 0.0f, T2MB, Q2B, CHB2;
  {
    MoninObukhov mb;
    mb.MOZ = 0.0f; mb.FM = 0.0f; mb.FH = 0.0f; mb.FM2 = 0.0f; mb.FH2 = 0.0f; mb.FV = 0.1f; mb.CM = CMB; mb.CH = CHB;
    mb.CH2 = 0.0f; mb.MOZSGN = 0;
    float WSTARb = 0.0f, Hb = 0.0f;
    const float ZPDb = ZPDG, Z0Mb = Z0MG, Z0Hb = Z0MG;
    if (ZLVL <= ZPDb) r.flags |= FLAG_ZLVL;                                               // func:3412-3415
    const float TMPCMb = m_logf((ZLVL - ZPDb) / Z0Mb), TMPCHb = m_logf((ZLVL - ZPDb) / Z0Hb);
    const float TMPCM2b = m_logf((2.0f + Z0Mb) / Z0Mb), TMPCH2b = m_logf((2.0f + Z0Hb) / Z0Hb);
    const float CIRb = EMG * SB;                                                          // func:3131-3132
    const float CGHb = 2.0f * DFtop / DZtop;
    float ESTGb = 0.0f, CSHb = 0.0f, CEVb = 0.0f, EHB = 0.0f;
#pragma unroll 1
    for (int iter = 1; iter <= 5; ++iter) {                                               // loop3, func:3134-3222
      NMP_PHASE_SYNC_AT(11);
      if (opt.sfc == 1) {
        sfcdif1(iter, f.SFCTMP, RHOAIR, Hb, QAIR, ZLVL, ZPDb, Z0Hb, UR, TMPCMb, TMPCHb, TMPCM2b, TMPCH2b, mb);
      }
      if (opt.sfc == 2) {                                                                 // func:3150-3164
        sfcdif2(iter, Z0Mb, TGB, THAIR, UR, T.gen[G_CZIL], ZLVL, mb.CM, mb.CH, mb.MOZ, WSTARb, mb.FV);
        mb.CH = mb.CH / UR;
        mb.CM = mb.CM / UR;
        if (c.snowh > 0.0f) {
          mb.CM = fmin2(0.01f, mb.CM);
          mb.CH = fmin2(0.01f, mb.CH);
        }
      }
#ifdef NMP_FAST_ALGEBRA
      EHB = fmin2(1.0f, mb.CH * UR);                                                      // = 1 / RAHB
      const float RAWB = 1.0f / EHB;
      float Tc = tdc(TGB);
      float DESTG;
      esat_sel(Tc, ESTGb, DESTG);
      CSHb = RHOAIR * CPAIR * EHB;
      CEVb = RHOAIR * CPAIR / (GAMMAG * (RSURF + RAWB));
#else
      const float RAHB = fmax2(1.0f, 1.0f / (mb.CH * UR));                                // func:3166-3172
      const float RAWB = RAHB;
      EHB = 1.0f / RAHB;
      float Tc = tdc(TGB);                                                                // func:3176-3184
      float DESTG;
      esat_sel(Tc, ESTGb, DESTG);
      CSHb = RHOAIR * CPAIR / RAHB;                                                       // func:3186-3187
      CEVb = RHOAIR * CPAIR / GAMMAG / (RSURF + RAWB);
#endif
      IRB = CIRb * pow4(TGB) - EMG * f.LWDN;                                              // func:3191-3194
      SHB = CSHb * (TGB - f.SFCTMP);
      EVB = CEVb * (ESTGb * RHSUR - EAIR);
      GHB = CGHb * (TGB - STCtop);
      const float B = SAG - IRB - SHB - EVB - GHB;                                        // func:3196-3198
      const float A = 4.0f * CIRb * cube(TGB) + CSHb + CEVb * DESTG + CGHb;
      const float DTG = B / A;
      IRB = IRB + 4.0f * CIRb * cube(TGB) * DTG;                                          // func:3200-3206
      SHB = SHB + CSHb * DTG;
      EVB = EVB + CEVb * DESTG * DTG;
      GHB = GHB + CGHb * DTG;
      TGB = TGB + DTG;
      Hb = CSHb * (TGB - f.SFCTMP);                                                       // func:3209
      Tc = tdc(TGB);                                                                      // func:3211-3218
      { float unused; esat_sel(Tc, ESTGb, unused); }
      QSFC = 0.622f * (ESTGb * RHSUR) / (PSFC - 0.378f * (ESTGb * RHSUR));
    }
    if (opt.stc == 1) {                                                                   // func:3225-3233
      if (c.snowh > 0.05f && TGB > TFRZ) {
        TGB = TFRZ;
        IRB = CIRb * pow4(TGB) - EMG * f.LWDN;
        SHB = CSHb * (TGB - f.SFCTMP);
        EVB = CEVb * (ESTGb * RHSUR - EAIR);
        GHB = SAG - (IRB + SHB + EVB);
      }
    }
    TAUXB = -RHOAIR * mb.CM * UR * f.UU;                                                  // func:3236-3237
    TAUYB = -RHOAIR * mb.CM * UR * f.VV;
    CHB2 = mb.FV * KARMAN / (m_logf((2.0f + Z0Hb) / Z0Hb) - mb.FH2);                      // func:3243-3252
    if (CHB2 < 1.0E-5f) {
      T2MB = TGB;
      Q2B = QSFC;
    } else {
      T2MB = TGB - SHB / (RHOAIR * CPAIR) * 1.0f / CHB2;
      Q2B = QSFC - EVB / (LATHEAG * RHOAIR) * (1.0f / CHB2 + RSURF);
    }
    if (c.lutyp == ISURBAN) Q2B = QSFC;
    CMB = mb.CM;
    CHB = EHB;                                                                            // func:3256
  }
  (void)TAUXV; (void)TAUYV; (void)TAUXB; (void)TAUYB;   // wind stress is computed by the reference but never returned

  NMP_PHASE_SYNC_AT(12);
  // ---- tile aggregation, func:1246-1308 ---------------------------------------------------------
  float FIRA, FSH, FGEV, SSOIL, FCEV, FCTR, TS;
  if (VEG && fveg > 0.0f) {
    FIRA = fveg * IRG + (1.0f - fveg) * IRB + IRC;
    FSH = fveg * SHG + (1.0f - fveg) * SHB + SHC;
    FGEV = fveg * EVG + (1.0f - fveg) * EVB;
    SSOIL = fveg * GHV + (1.0f - fveg) * GHB;
    FCEV = EVC;
    FCTR = TR;
    c.TG = fveg * TGV + (1.0f - fveg) * TGB;
    TS = fveg * c.TV + (1.0f - fveg) * TGB;
    c.CM = fveg * CMV + (1.0f - fveg) * CMB;
    c.CH = fveg * CHV + (1.0f - fveg) * CHB;
  } else {
    FIRA = IRB; FSH = SHB; FGEV = EVB; SSOIL = GHB;
    c.TG = TGB;
    FCEV = 0.0f; FCTR = 0.0f;
    TS = c.TG;
    c.CM = CMB; c.CH = CHB;
    RSSUN = 0.0f; RSSHA = 0.0f;
    TGV = TGB;
    CHV = CHB;
  }
  (void)TS;
  c.QSFC = QSFC;
  const float FIRE = f.LWDN + FIRA;                                                       // func:1284-1292
  if (FIRE <= 0.0f) r.flags |= FLAG_FIRE;
  const float EMISSI = fveg * (EMG * (1.0f - EMV) + EMV + EMV * (1.0f - EMV) * (1.0f - EMG)) + (1.0f - fveg) * EMG;
  const float TRAD = m_powf((FIRE - (1.0f - EMISSI) * f.LWDN) / (EMISSI * SB), 0.25f);      // func:1302
  r.APAR = PARSUN * LAISUN + PARSHA * LAISHA;                                             // func:1307-1308
  r.PSN = PSNSUN * LAISUN + PSNSHA * LAISHA;
  r.FIRA = FIRA; r.FSH = FSH; r.FGEV = FGEV; r.SSOIL = SSOIL; r.FCEV = FCEV; r.FCTR = FCTR;
  r.CH_out = c.CH;
  r.Q2B = Q2B;

  NMP_PHASE_SYNC_AT(13);
  // ---- thermoprop, second half: the remaining conductivities, heat capacities (csnow 1476-1490, func:1398-1403, lakes
  // 1420-1430) and FACT (1434-1436)
  layer_thickness_again(c, DZSNSO);
  float DF[7], HCPCT[7], FACT[7];
  static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
    HCPCT[IX7(k)] = 0.0f; DF[IX7(k)] = 0.0f;
    if (k >= ISNOW + 1) {
      const float dz = DZSNSO[IX7(k)];
      const float SNICEV = fmin2(1.0f, c.SNICE[IX3(k)] / (dz * DENICE));                  // func:1476-1478
      const float epore = 1.0f - SNICEV;
      const float SNLIQV = fmin2(epore, c.SNLIQ[IX3(k)] / (dz * DENWAT));
      const float BDSNOI = (c.SNICE[IX3(k)] + c.SNLIQ[IX3(k)]) / dz;                      // func:1482-1490
      HCPCT[IX7(k)] = CICE * SNICEV + CWAT * SNLIQV;
      DF[IX7(k)] = 3.2217E-6f * (BDSNOI * BDSNOI);
    }
  });
  DF[IX7(1)] = DF1;
  static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
    if (k >= 2) {
      NMP_PHASE_SYNC_AT(4);
      DF[IX7(k)] = tdfcnd(SMCMAX, SP[P_QUARTZ], c.SMC[IX4(k)], c.soilwat[IX4(k)]);
      if (c.lutyp == ISURBAN) DF[IX7(k)] = 3.24f;                                         // func:1405-1409
    }
    const float soilice = c.SMC[IX4(k)] - c.soilwat[IX4(k)];
    HCPCT[IX7(k)] = c.soilwat[IX4(k)] * CWAT + (1.0f - SMCMAX) * T.gen[G_CSOIL] + (SMCMAX - c.SMC[IX4(k)]) * CPAIR +
                    soilice * CICE;
    if (c.IST == 2) {                                                                     // func:1420-1430
      if (c.STC[IX7(k)] > TFRZ) { HCPCT[IX7(k)] = CWAT; if (k >= 2) DF[IX7(k)] = TKWAT; }
      else { HCPCT[IX7(k)] = CICE; if (k >= 2) DF[IX7(k)] = TKICE; }
    }
  });
  static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
    FACT[IX7(k)] = (k >= ISNOW + 1) ? DT / (HCPCT[IX7(k)] * DZSNSO[IX7(k)]) : 0.0f;
  });
  // ---- tsnosoi: hrt + hstep + rosr12, func:3987-4288 ---------------------------------------------
  {
    const float ZBOTSNO = T.gen[G_ZBOT] - c.snowh;                                        // func:4039
    float AI[7], BI[7], CI[7], RHS[7];
    float dtsdz_prev = 0.0f, ddz_prev = 0.0f;
    static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;                                                   // func:4139-4186 (two passes fused)
      AI[IX7(k)] = 0.0f; BI[IX7(k)] = 0.0f; CI[IX7(k)] = 0.0f; RHS[IX7(k)] = 0.0f;
      if (k >= ISNOW + 1) {
        float DENOM, DDZ = 0.0f, DTSDZ = 0.0f, EFLUX;
        if (k == ISNOW + 1) {
          DENOM = -c.ZSNSO[IX7(k)] * HCPCT[IX7(k)];
          const float TEMP1 = -c.ZSNSO[IX7(k < NSOIL ? k + 1 : k)];
          DDZ = 2.0f / TEMP1;
          DTSDZ = 2.0f * (c.STC[IX7(k)] - c.STC[IX7(k < NSOIL ? k + 1 : k)]) / TEMP1;
          EFLUX = DF[IX7(k)] * DTSDZ - SSOIL - 0.0f;
          AI[IX7(k)] = 0.0f;
          CI[IX7(k)] = -DF[IX7(k)] * DDZ / DENOM;
          if (opt.stc == 1) BI[IX7(k)] = -CI[IX7(k)];
          if (opt.stc == 2) BI[IX7(k)] = -CI[IX7(k)] + DF[IX7(k)] / (0.5f * c.ZSNSO[IX7(k)] * c.ZSNSO[IX7(k)] * HCPCT[IX7(k)]);
        } else if (k < NSOIL) {
          DENOM = (c.ZSNSO[IX7(k > -2 ? k - 1 : k)] - c.ZSNSO[IX7(k)]) * HCPCT[IX7(k)];
          const float TEMP1 = c.ZSNSO[IX7(k > -2 ? k - 1 : k)] - c.ZSNSO[IX7(k < NSOIL ? k + 1 : k)];
          DDZ = 2.0f / TEMP1;
          DTSDZ = 2.0f * (c.STC[IX7(k)] - c.STC[IX7(k < NSOIL ? k + 1 : k)]) / TEMP1;
          EFLUX = (DF[IX7(k)] * DTSDZ - DF[IX7(k > -2 ? k - 1 : k)] * dtsdz_prev) - 0.0f;
          AI[IX7(k)] = -DF[IX7(k > -2 ? k - 1 : k)] * ddz_prev / DENOM;
          CI[IX7(k)] = -DF[IX7(k)] * DDZ / DENOM;
          BI[IX7(k)] = -(AI[IX7(k)] + CI[IX7(k)]);
        } else {
          DENOM = (c.ZSNSO[IX7(k > -2 ? k - 1 : k)] - c.ZSNSO[IX7(k)]) * HCPCT[IX7(k)];
          float BOTFLX = 0.0f;
          if (opt.tbot == 2) {
            DTSDZ = (c.STC[IX7(k)] - c.TBOT) / (0.5f * (c.ZSNSO[IX7(k > -2 ? k - 1 : k)] + c.ZSNSO[IX7(k)]) - ZBOTSNO);
            BOTFLX = -DF[IX7(k)] * DTSDZ;
          }
          EFLUX = (-BOTFLX - DF[IX7(k > -2 ? k - 1 : k)] * dtsdz_prev) - 0.0f;
          AI[IX7(k)] = -DF[IX7(k > -2 ? k - 1 : k)] * ddz_prev / DENOM;
          CI[IX7(k)] = 0.0f;
          BI[IX7(k)] = -(AI[IX7(k)] + CI[IX7(k)]);
        }
        RHS[IX7(k)] = EFLUX / (-DENOM);
        dtsdz_prev = DTSDZ;
        ddz_prev = DDZ;
      }
    });
    // hstep (func:4213-4235) + rosr12 (func:4268-4287): Thomas algorithm on layers ISNOW+1..4
    float P[7], DELTA[7];
    static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      P[IX7(k)] = 0.0f; DELTA[IX7(k)] = 0.0f;
      if (k >= ISNOW + 1) {
        const float rhs = RHS[IX7(k)] * DT, a = AI[IX7(k)] * DT, b = 1.0f + BI[IX7(k)] * DT;
        const float cc = (k == NSOIL) ? 0.0f : CI[IX7(k)] * DT;
        if (k == ISNOW + 1) {
          P[IX7(k)] = -cc / b;
          DELTA[IX7(k)] = rhs / b;
        } else {
          const float pm = P[IX7(k > -2 ? k - 1 : k)], dm = DELTA[IX7(k > -2 ? k - 1 : k)];
          P[IX7(k)] = -cc * (1.0f / (b + a * pm));
          DELTA[IX7(k)] = (rhs - a * dm) * (1.0f / (b + a * pm));
        }
      }
    });
    P[IX7(NSOIL)] = DELTA[IX7(NSOIL)];
    static_for_rev<NSOIL - 1, -2>([&](auto k_) { constexpr int k = decltype(k_)::value;
      if (k >= ISNOW + 1) P[IX7(k)] = P[IX7(k)] * P[IX7(k + 1)] + DELTA[IX7(k)];
    });
    static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      if (k >= ISNOW + 1) c.STC[IX7(k)] = c.STC[IX7(k)] + P[IX7(k)];
    });
  }

  if (opt.stc == 2) {                                                                     // func:1318-1330
    if (c.snowh > 0.05f && c.TG > TFRZ) {
      TGV = TFRZ;
      TGB = TFRZ;
      c.TG = (VEG && fveg > 0.0f) ? fveg * TGV + (1.0f - fveg) * TGB : TGB;
    }
  }

  NMP_PHASE_SYNC_AT(14);
  // ---- phasechange, func:4291-4491 (layers are independent: one pass, per-layer scalars) ---------
  {
    float QMELT = 0.0f, PONDING = 0.0f;
    static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      if (k <= 0) r.IMELT_snow[IX3(k <= 0 ? k : 0)] = 0;
      if (k >= ISNOW + 1) {
        float MICE, MLIQ, SUPERCOOL = 0.0f;
        if (k <= 0) {
          MICE = c.SNICE[IX3(k <= 0 ? k : 0)];
          MLIQ = c.SNLIQ[IX3(k <= 0 ? k : 0)];
        } else {
          MLIQ = c.soilwat[IX4(k >= 1 ? k : 1)] * DZSNSO[IX7(k)] * 1000.0f;               // func:4360-4361
          MICE = (c.SMC[IX4(k >= 1 ? k : 1)] - c.soilwat[IX4(k >= 1 ? k : 1)]) * DZSNSO[IX7(k)] * 1000.0f;
          if (c.IST == 1) {                                                               // func:4373-4387
            if (opt.frz == 1) {
              if (c.STC[IX7(k)] < TFRZ) {
                const float SMP = HFUS * (TFRZ - c.STC[IX7(k)]) / (GRAV * c.STC[IX7(k)]);
                SUPERCOOL = SMCMAX * m_powf(SMP / PSISAT, -1.0f / BEXP);
                SUPERCOOL = SUPERCOOL * DZSNSO[IX7(k)] * 1000.0f;
              }
            }
            if (opt.frz == 2) {                                                           // frh2o, func:4494-4598
              const float TK = c.STC[IX7(k)], smc = c.SMC[IX4(k >= 1 ? k : 1)], sh2o = c.soilwat[IX4(k >= 1 ? k : 1)];
              float FREE = 0.0f;
              float BX = BEXP;
              if (BEXP > 5.5f) BX = 5.5f;
              if (TK > (TFRZ - 1.0E-3f)) {
                FREE = smc;
              } else {
                int NLOG = 0, KCOUNT = 0;
                float SWL = smc - sh2o;
                if (SWL > (smc - 0.02f)) SWL = smc - 0.02f;
                if (SWL < 0.0f) SWL = 0.0f;
#pragma unroll 1
                while ((NLOG < 10) && (KCOUNT == 0)) {
                  NLOG = NLOG + 1;
                  const float DFr = m_logf((PSISAT * GRAV / HFUS) * (sq(1.0f + 8.0f * SWL)) * m_powf(SMCMAX / (smc - SWL), BX)) -
                                    m_logf(-(TK - TFRZ) / TK);
                  const float DENOM = 2.0f * 8.0f / (1.0f + 8.0f * SWL) + BX / (smc - SWL);
                  float SWLK = SWL - DFr / DENOM;
                  if (SWLK > (smc - 0.02f)) SWLK = smc - 0.02f;
                  if (SWLK < 0.0f) SWLK = 0.0f;
                  const float DSWL = fabsf(SWLK - SWL);
                  SWL = SWLK;
                  if (DSWL <= 0.005f) KCOUNT = KCOUNT + 1;
                }
                FREE = smc - SWL;
                if (KCOUNT == 0) {
                  float FK = (m_powf((HFUS / (GRAV * (-PSISAT))) * ((TK - TFRZ) / TK), -1.0f / BX)) * SMCMAX;
                  if (FK < 0.02f) FK = 0.02f;
                  FREE = fmin2(FK, smc);
                }
              }
              SUPERCOOL = FREE * DZSNSO[IX7(k)] * 1000.0f;
            }
          }
        }
        const float WICE0 = MICE, WMASS0 = MICE + MLIQ;
        int IMELT = 0;                                                                    // func:4389-4403
        if (MICE > 0.0f && c.STC[IX7(k)] >= TFRZ) IMELT = 1;
        if (MLIQ > SUPERCOOL && c.STC[IX7(k)] < TFRZ) IMELT = 2;
        if (k == 1 && ISNOW == 0 && c.sneqv > 0.0f && c.STC[IX7(k)] >= TFRZ) IMELT = 1;
        float HM = 0.0f;                                                                  // func:4406-4421
        if (IMELT > 0) {
          HM = (c.STC[IX7(k)] - TFRZ) / FACT[IX7(k)];
          c.STC[IX7(k)] = TFRZ;
        }
        if (IMELT == 1 && HM < 0.0f) { HM = 0.0f; IMELT = 0; }
        if (IMELT == 2 && HM > 0.0f) { HM = 0.0f; IMELT = 0; }
        float XM = HM * DT / HFUS;
        if (k == 1 && ISNOW == 0 && c.sneqv > 0.0f && XM > 0.0f) {                        // func:4424-4440
          const float TEMP1 = c.sneqv;
          c.sneqv = fmax2(0.0f, TEMP1 - XM);
          const float PROPOR = c.sneqv / TEMP1;
          c.snowh = fmax2(0.0f, PROPOR * c.snowh);
          const float HEATR = HM - HFUS * (TEMP1 - c.sneqv) / DT;
          if (HEATR > 0.0f) { XM = HEATR * DT / HFUS; HM = HEATR; }
          else { XM = 0.0f; HM = 0.0f; }
          QMELT = fmax2(0.0f, (TEMP1 - c.sneqv)) / DT;
          PONDING = TEMP1 - c.sneqv;
        }
        if (IMELT > 0 && fabsf(HM) > 0.0f) {                                              // func:4443-4479
          float HEATR = 0.0f;
          if (XM > 0.0f) {
            MICE = fmax2(0.0f, WICE0 - XM);
            HEATR = HM - HFUS * (WICE0 - MICE) / DT;
          } else if (XM < 0.0f) {
            if (k <= 0) {
              MICE = fmin2(WMASS0, WICE0 - XM);
            } else {
              if (WMASS0 < SUPERCOOL) MICE = 0.0f;
              else { MICE = fmin2(WMASS0 - SUPERCOOL, WICE0 - XM); MICE = fmax2(MICE, 0.0f); }
            }
            HEATR = HM - HFUS * (WICE0 - MICE) / DT;
          }
          MLIQ = fmax2(0.0f, WMASS0 - MICE);
          if (fabsf(HEATR) > 0.0f) {
            c.STC[IX7(k)] = c.STC[IX7(k)] + FACT[IX7(k)] * HEATR;
            if (k <= 0) { if (MLIQ * MICE > 0.0f) c.STC[IX7(k)] = TFRZ; }
          }
          if (k < 1) QMELT = QMELT + fmax2(0.0f, (WICE0 - MICE)) / DT;
        }
        if (k <= 0) {                                                                     // func:4481-4489
          c.SNLIQ[IX3(k <= 0 ? k : 0)] = MLIQ;
          c.SNICE[IX3(k <= 0 ? k : 0)] = MICE;
          r.IMELT_snow[IX3(k <= 0 ? k : 0)] = IMELT;
        } else {
          c.soilwat[IX4(k >= 1 ? k : 1)] = div_rn(MLIQ, 1000.0f * DZSNSO[IX7(k)]);
          c.SMC[IX4(k >= 1 ? k : 1)] = div_rn(MLIQ + MICE, 1000.0f * DZSNSO[IX7(k)]);
        }
      }
    });
    r.QMELT = QMELT;
    r.PONDING = PONDING;
  }

  if (DIAG) {
    dg.put(D_FSA, FSA); dg.put(D_FSR, FSR); dg.put(D_FIRA, FIRA); dg.put(D_FSH, FSH); dg.put(D_SSOIL, SSOIL);
    dg.put(D_FCEV, FCEV); dg.put(D_FGEV, FGEV); dg.put(D_FCTR, FCTR); dg.put(D_TRAD, TRAD); dg.put(D_TGB, TGB);
    dg.put(D_TGV, TGV); dg.put(D_T2MV, T2MV); dg.put(D_T2MB, T2MB); dg.put(D_Q2V, Q2V); dg.put(D_APAR, r.APAR);
    dg.put(D_PSN, r.PSN); dg.put(D_SAV, SAV); dg.put(D_SAG, SAG); dg.put(D_PONDING, r.PONDING);
    dg.put(D_RSSUN, RSSUN); dg.put(D_RSSHA, RSSHA); dg.put(D_CHV, CHV); dg.put(D_CHB, CHB); dg.put(D_EMISSI, EMISSI);
    dg.put(D_SHG, SHG); dg.put(D_SHC, SHC); dg.put(D_SHB, SHB); dg.put(D_EVG, EVG); dg.put(D_EVB, EVB);
    dg.put(D_GHV, GHV); dg.put(D_GHB, GHB); dg.put(D_IRG, IRG); dg.put(D_IRC, IRC); dg.put(D_IRB, IRB);
    dg.put(D_TR, TR); dg.put(D_EVC, EVC); dg.put(D_CHLEAF, CHLEAF); dg.put(D_CHUC, CHUC); dg.put(D_CHV2, CHV2);
    dg.put(D_CHB2, CHB2);
  }
}

}  // namespace nmp
