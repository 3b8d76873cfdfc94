// noahmp_b200 -- water half of the column step: canopy water, snowpack, soil moisture, groundwater,
// carbon.  Reference: core/module_noahmp_func.f90:4601-7025 ("func:").
//
// Register-resident like energy.cuh.  The only dynamically indexed code is the snow-layer
// combine/divide pair, whose loop bounds and merge targets depend on the data; it runs on a small
// scratch copy (SnowPack) inside two non-inlined functions so the rest of the column stays in registers.
#pragma once
#include "energy.cuh"

namespace nmp {

struct SnowPack {   // snow layers -2..0 at [k+2], plus the scalars combine/divide update
  float dz[3], ice[3], liq[3], t[3];
  int isnow;
  float soilwat1, soilice1, snowh, sneqv, ponding1, ponding2;
};

// ---- func:5536-5577 ----------------------------------------------------------------------------------
__device__ __forceinline__ void combo(float& DZ, float& WLIQ, float& WICE, float& T, float DZ2, float WLIQ2,
                                      float WICE2, float T2) {
  const float DZC = DZ + DZ2, WICEC = WICE + WICE2, WLIQC = WLIQ + WLIQ2;
  const float H = (CICE * WICE + CWAT * WLIQ) * (T - TFRZ) + HFUS * WLIQ;
  const float H2 = (CICE * WICE2 + CWAT * WLIQ2) * (T2 - TFRZ) + HFUS * WLIQ2;
  const float HC = H + H2;
  float TC;
  if (HC < 0.0f) TC = TFRZ + HC / (CICE * WICEC + CWAT * WLIQC);
  else if (HC <= HFUS * WLIQC) TC = TFRZ;
  else TC = TFRZ + (HC - HFUS * WLIQC) / (CICE * WICEC + CWAT * WLIQC);
  DZ = DZC; WICE = WICEC; WLIQ = WLIQC; T = TC;
}

// ---- func:5236-5413 ----------------------------------------------------------------------------------
__device__ __forceinline__ void combine_body(SnowPack& s, float dzsoil1, int& ISNOW, float& soilwat1,
                                             float& soilice1, float& snowh, float& sneqv, float& PONDING1,
                                             float& PONDING2) {
#define L(k) NMP_CHK((k) + 2, 0, 2, 1)
  int ISNOW_OLD = ISNOW;

  for (int J = ISNOW_OLD + 1; J <= 0; ++J) {                                              // func:5276-5318
    if (s.ice[L(J)] <= 0.1f) {
      if (J != 0) {
        s.liq[L(J + 1)] = s.liq[L(J + 1)] + s.liq[L(J)];
        s.ice[L(J + 1)] = s.ice[L(J + 1)] + s.ice[L(J)];
      } else {
        if (ISNOW_OLD < -1) {
          s.liq[L(J - 1)] = s.liq[L(J - 1)] + s.liq[L(J)];
          s.ice[L(J - 1)] = s.ice[L(J - 1)] + s.ice[L(J)];
        } else {
          if (s.ice[L(J)] >= 0.0f) {
            PONDING1 = s.liq[L(J)];
            sneqv = s.ice[L(J)];
            snowh = s.dz[L(J)];
          } else {
            PONDING1 = s.liq[L(J)] + s.ice[L(J)];
            if (PONDING1 < 0.0f) {
              soilice1 = fmax2(0.0f, soilice1 + PONDING1 / (dzsoil1 * 1000.0f));
              PONDING1 = 0.0f;
            }
            sneqv = 0.0f;
            snowh = 0.0f;
          }
          s.liq[L(J)] = 0.0f;
          s.ice[L(J)] = 0.0f;
          s.dz[L(J)] = 0.0f;
        }
      }
      if (J > ISNOW + 1 && ISNOW < -1) {                                                  // func:5308-5315
        for (int I = J; I >= ISNOW + 2; --I) {
          s.t[L(I)] = s.t[L(I - 1)];
          s.liq[L(I)] = s.liq[L(I - 1)];
          s.ice[L(I)] = s.ice[L(I - 1)];
          s.dz[L(I)] = s.dz[L(I - 1)];
        }
      }
      ISNOW = ISNOW + 1;
    }
  }
  if (soilice1 < 0.0f) {                                                                  // func:5322-5325
    soilwat1 = soilwat1 + soilice1;
    soilice1 = 0.0f;
  }
  if (ISNOW == 0) return;                                                                 // func:5327
  sneqv = 0.0f;                                                                           // func:5329-5339
  snowh = 0.0f;
  float ZWICE = 0.0f, ZWLIQ = 0.0f;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    sneqv = sneqv + s.ice[L(J)] + s.liq[L(J)];
    snowh = snowh + s.dz[L(J)];
    ZWICE = ZWICE + s.ice[L(J)];
    ZWLIQ = ZWLIQ + s.liq[L(J)];
  }
  if (snowh < 0.025f && ISNOW < 0) {                                                      // func:5344-5350
    ISNOW = 0;
    sneqv = ZWICE;
    PONDING2 = ZWLIQ;
    if (sneqv <= 0.0f) snowh = 0.0f;
  }
  if (ISNOW < -1) {                                                                       // func:5361-5411
    ISNOW_OLD = ISNOW;
    int MSSI = 1;
    for (int I = ISNOW_OLD + 1; I <= 0; ++I) {
      const float dzmin = (MSSI == 3) ? 0.1f : 0.025f;                                    // DZMIN = (0.025, 0.025, 0.1)
      if (s.dz[L(I)] < dzmin) {
        int NEIBOR;
        if (I == ISNOW + 1) NEIBOR = I + 1;
        else if (I == 0) NEIBOR = I - 1;
        else {
          NEIBOR = I + 1;
          if ((s.dz[L(I - 1)] + s.dz[L(I)]) < (s.dz[L(I + 1)] + s.dz[L(I)])) NEIBOR = I - 1;
        }
        int J, LL;
        if (NEIBOR > I) { J = NEIBOR; LL = I; } else { J = I; LL = NEIBOR; }
        combo(s.dz[L(J)], s.liq[L(J)], s.ice[L(J)], s.t[L(J)], s.dz[L(LL)], s.liq[L(LL)], s.ice[L(LL)], s.t[L(LL)]);
        if (J - 1 > ISNOW + 1) {                                                          // func:5391-5398
          for (int K = J - 1; K >= ISNOW + 2; --K) {
            s.t[L(K)] = s.t[L(K - 1)];
            s.ice[L(K)] = s.ice[L(K - 1)];
            s.liq[L(K)] = s.liq[L(K - 1)];
            s.dz[L(K)] = s.dz[L(K - 1)];
          }
        }
        ISNOW = ISNOW + 1;
        if (ISNOW >= -1) break;
      } else {
        MSSI = MSSI + 1;
      }
    }
  }
#undef L
}
// Non-inlined wrapper: everything combine touches travels in the SnowPack scratch, so no register of the
// caller has its address taken.
__device__ __noinline__ void combine(SnowPack& s, float dzsoil1) {
  int ISNOW = s.isnow;
  float soilwat1 = s.soilwat1, soilice1 = s.soilice1, snowh = s.snowh, sneqv = s.sneqv;
  float PONDING1 = s.ponding1, PONDING2 = s.ponding2;
  combine_body(s, dzsoil1, ISNOW, soilwat1, soilice1, snowh, sneqv, PONDING1, PONDING2);
  s.isnow = ISNOW; s.soilwat1 = soilwat1; s.soilice1 = soilice1; s.snowh = snowh; s.sneqv = sneqv;
  s.ponding1 = PONDING1; s.ponding2 = PONDING2;
}

// ---- func:5416-5533 ----------------------------------------------------------------------------------
__device__ __noinline__ void divide(SnowPack& s) {
  int ISNOW = s.isnow;
  float DZ[3] = {0.0f, 0.0f, 0.0f}, SWICE[3] = {0.0f, 0.0f, 0.0f}, SWLIQ[3] = {0.0f, 0.0f, 0.0f}, TSNO[3] = {0.0f, 0.0f, 0.0f};
  int MSNO = -ISNOW;
  for (int J = 1; J <= NSNOW; ++J) {                                                      // func:5443-5450
    if (J <= MSNO) {
      DZ[NMP_CHK(J - 1, 0, 2, 2)] = s.dz[NMP_CHK(J + ISNOW + 2, 0, 2, 2)];
      SWICE[NMP_CHK(J - 1, 0, 2, 2)] = s.ice[NMP_CHK(J + ISNOW + 2, 0, 2, 2)];
      SWLIQ[NMP_CHK(J - 1, 0, 2, 2)] = s.liq[NMP_CHK(J + ISNOW + 2, 0, 2, 2)];
      TSNO[NMP_CHK(J - 1, 0, 2, 2)] = s.t[NMP_CHK(J + ISNOW + 2, 0, 2, 2)];
    }
  }
  if (MSNO == 1) {                                                                        // func:5454-5466
    if (DZ[0] > 0.05f) {
      MSNO = 2;
      DZ[0] = DZ[0] / 2.0f; SWICE[0] = SWICE[0] / 2.0f; SWLIQ[0] = SWLIQ[0] / 2.0f;
      DZ[1] = DZ[0]; SWICE[1] = SWICE[0]; SWLIQ[1] = SWLIQ[0]; TSNO[1] = TSNO[0];
    }
  }
  if (MSNO > 1) {                                                                         // func:5468-5502
    if (DZ[0] > 0.05f) {
      const float DRR = DZ[0] - 0.05f;
      float PROPOR = DRR / DZ[0];
      const float ZWICE = PROPOR * SWICE[0], ZWLIQ = PROPOR * SWLIQ[0];
      PROPOR = 0.05f / DZ[0];
      SWICE[0] = PROPOR * SWICE[0];
      SWLIQ[0] = PROPOR * SWLIQ[0];
      DZ[0] = 0.05f;
      combo(DZ[1], SWLIQ[1], SWICE[1], TSNO[1], DRR, ZWLIQ, ZWICE, TSNO[0]);
      if (MSNO <= 2 && DZ[1] > 0.20f) {                                                   // func:5483-5499
        MSNO = 3;
        const float DTDZ = (TSNO[0] - TSNO[1]) / ((DZ[0] + DZ[1]) / 2.0f);
        DZ[1] = DZ[1] / 2.0f; SWICE[1] = SWICE[1] / 2.0f; SWLIQ[1] = SWLIQ[1] / 2.0f;
        DZ[2] = DZ[1]; SWICE[2] = SWICE[1]; SWLIQ[2] = SWLIQ[1];
        TSNO[2] = TSNO[1] - DTDZ * DZ[1] / 2.0f;
        if (TSNO[2] >= TFRZ) TSNO[2] = TSNO[1];
        else TSNO[1] = TSNO[1] + DTDZ * DZ[1] / 2.0f;
      }
    }
  }
  if (MSNO > 2) {                                                                         // func:5504-5517
    if (DZ[1] > 0.2f) {
      const float DRR = DZ[1] - 0.2f;
      float PROPOR = DRR / DZ[1];
      const float ZWICE = PROPOR * SWICE[1], ZWLIQ = PROPOR * SWLIQ[1];
      PROPOR = 0.2f / DZ[1];
      SWICE[1] = PROPOR * SWICE[1];
      SWLIQ[1] = PROPOR * SWLIQ[1];
      DZ[1] = 0.2f;
      combo(DZ[2], SWLIQ[2], SWICE[2], TSNO[2], DRR, ZWLIQ, ZWICE, TSNO[1]);
    }
  }
  ISNOW = -MSNO;                                                                          // func:5519-5526
  s.isnow = ISNOW;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    s.dz[NMP_CHK(J + 2, 0, 2, 2)] = DZ[NMP_CHK(J - ISNOW - 1, 0, 2, 2)];
    s.ice[NMP_CHK(J + 2, 0, 2, 2)] = SWICE[NMP_CHK(J - ISNOW - 1, 0, 2, 2)];
    s.liq[NMP_CHK(J + 2, 0, 2, 2)] = SWLIQ[NMP_CHK(J - ISNOW - 1, 0, 2, 2)];
    s.t[NMP_CHK(J + 2, 0, 2, 2)] = TSNO[NMP_CHK(J - ISNOW - 1, 0, 2, 2)];
  }
}

__device__ __forceinline__ void pack_snow(const Column& c, const float (&DZSNSO)[7], SnowPack& s) {
  static_for<0, (3) - 1>([&](auto i_) { constexpr int i = decltype(i_)::value; s.dz[i] = DZSNSO[i]; s.ice[i] = c.SNICE[i]; s.liq[i] = c.SNLIQ[i]; s.t[i] = c.STC[i]; });
  s.isnow = c.ISNOW; s.soilwat1 = c.soilwat[0]; s.snowh = c.snowh; s.sneqv = c.sneqv;
}
__device__ __forceinline__ void unpack_snow(const SnowPack& s, Column& c, float (&DZSNSO)[7]) {
  static_for<0, (3) - 1>([&](auto i_) { constexpr int i = decltype(i_)::value; DZSNSO[i] = s.dz[i]; c.SNICE[i] = s.ice[i]; c.SNLIQ[i] = s.liq[i]; c.STC[i] = s.t[i]; });
  c.ISNOW = s.isnow; c.soilwat[0] = s.soilwat1; c.snowh = s.snowh; c.sneqv = s.sneqv;
}

// ---- func:6386-6455 ----------------------------------------------------------------------------------
__device__ __forceinline__ void wdfcnd1(float smcmax, float bexp, float dwsat, float dksat, float SMC, float FCR,
                                        float& WDF, float& WCND) {
  const float FACTR = fmax2(0.01f, SMC / smcmax);
  WDF = dwsat * m_soilpowf(FACTR, bexp + 2.0f);
  WDF = WDF * (1.0f - FCR);
  WCND = dksat * m_soilpowf(FACTR, 2.0f * bexp + 3.0f);
  WCND = WCND * (1.0f - FCR);
}
__device__ __forceinline__ void wdfcnd2(float smcmax, float bexp, float dwsat, float dksat, float SMC,
                                        float soilice, float& WDF, float& WCND) {
  const float FACTR = fmax2(0.01f, SMC / smcmax);
  const float EXPON = bexp + 2.0f;
  WDF = dwsat * m_soilpowf(FACTR, EXPON);
  if (soilice > 0.0f) {
    const float VKWGT = 1.0f / (1.0f + m_soilpowf(500.0f * soilice, 3.0f));
    WDF = VKWGT * WDF + (1.0f - VKWGT) * dwsat * m_soilpowf(0.2f / smcmax, EXPON);
  }
  WCND = dksat * m_soilpowf(FACTR, 2.0f * bexp + 3.0f);
}

struct WaterResult {
  float ECAN, ETRAN, runsrf, runsub, QSNBOT, PONDING1, PONDING2, FPICE;
  int niter_soil;
};

// =====================================================================================================
// water: func:4601-4804 with canwater (4807-5046), snowwater (5049-5174: snowfall, compact, combine,
// divide, snowh2o), soilh2o (5822-6048: zwteq, infil, srt, sstep) and groundwater (6458-6639).
// =====================================================================================================
template <class Opt>
__device__ __forceinline__ void water(const Opt& opt, const Tables& T, const RunConst& rc, Column& c,
                                      const Forcing& f, const EnergyResult& e, float QPRECC, float QPRECL,
                                      float elai, float esai, float fveg, float QVAP, float QDEW,
                                      const float (&FICEOLD)[3], float (&soilice)[4], float (&DZSNSO)[7],
                                      WaterResult& w) {
  const float* VG = &T.veg[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGF];
  const float* SP = &T.soil[NMP_CHK(c.sltyp - 1, 0, MAXSOILTYP - 1, 4) * NSOILP];
  const float DT = rc.dt;
  const float SMCMAX = SP[P_SMCMAX], BEXP = SP[P_BEXP];
  const int ISURBAN = T.flags[0];
  float SNOFLOW = 0.0f;
  w.runsub = 0.0f;
  w.runsrf = 0.0f;
  w.niter_soil = 0;

  // ---- canwater, func:4807-5046 ---------------------------------------------------------------
  float QRAIN, QSNOW, snowhin;
  {
    float FP = 0.0f, FPICE = 0.0f;
    if (opt.snf == 1) {                                                                   // func:4893-4905
      if (f.SFCTMP > TFRZ + 2.5f) FPICE = 0.0f;
      else if (f.SFCTMP <= TFRZ + 0.5f) FPICE = 1.0f;
      else if (f.SFCTMP <= TFRZ + 2.0f) FPICE = 1.0f - (-54.632f + 0.2f * f.SFCTMP);
      else FPICE = 0.6f;
    }
    if (opt.snf == 2) FPICE = (f.SFCTMP >= TFRZ + 2.2f) ? 0.0f : 1.0f;                    // func:4907-4913
    if (opt.snf == 3) FPICE = (f.SFCTMP >= TFRZ) ? 0.0f : 1.0f;                           // func:4915-4921
    w.FPICE = FPICE;
    const float BDFALL = fmin2(120.0f, 67.92f + 51.25f * m_expf((f.SFCTMP - TFRZ) / 2.59f));  // func:4926
    const float RAIN = (QPRECC + QPRECL) * (1.0f - FPICE);
    const float SNOW = (QPRECC + QPRECL) * FPICE;
    if (QPRECC + QPRECL > 0.0f) FP = (QPRECC + QPRECL) / (10.0f * QPRECC + QPRECL);       // func:4933-4934
    const float MAXLIQ = VG[V_CANWMXP] * (elai + esai);                                   // func:4939
    float QINTR, QDRIPR, QTHROR, QINTS, QDRIPS, QTHROS;
    if ((elai + esai) > 0.0f) {                                                           // func:4943-4953
      QINTR = fveg * RAIN * FP;
      QINTR = fmin2(QINTR, (MAXLIQ - c.CANLIQ) / DT * (1.0f - m_expf(-RAIN * DT / MAXLIQ)));
      QINTR = fmax2(QINTR, 0.0f);
      QDRIPR = fveg * RAIN - QINTR;
      QTHROR = (1.0f - fveg) * RAIN;
    } else {
      QINTR = 0.0f; QDRIPR = 0.0f; QTHROR = RAIN;
    }
    float ETRAN, QEVAC, QDEWC, QSUBC, QFROC;
    if (!e.frozen_canopy) {                                                               // func:4956-4968
      ETRAN = fmax2(e.FCTR / HVAP, 0.0f);
      QEVAC = fmax2(e.FCEV / HVAP, 0.0f);
      QDEWC = fabsf(fmin2(e.FCEV / HVAP, 0.0f));
      QSUBC = 0.0f; QFROC = 0.0f;
    } else {
      ETRAN = fmax2(e.FCTR / HSUB, 0.0f);
      QEVAC = 0.0f; QDEWC = 0.0f;
      QSUBC = fmax2(e.FCEV / HSUB, 0.0f);
      QFROC = fabsf(fmin2(e.FCEV / HSUB, 0.0f));
    }
    QEVAC = fmin2(c.CANLIQ / DT, QEVAC);                                                  // func:4973-4975
    c.CANLIQ = fmax2(0.0f, c.CANLIQ + (QINTR + QDEWC - QEVAC) * DT);
    if (c.CANLIQ <= 1.0E-6f) c.CANLIQ = 0.0f;
    const float MAXSNO = 6.6f * (0.27f + 46.0f / BDFALL) * (elai + esai);                 // func:4978
    if ((elai + esai) > 0.0f) {                                                           // func:4980-4992
      QINTS = fveg * SNOW * FP;
      QINTS = fmin2(QINTS, (MAXSNO - c.CANICE) / DT * (1.0f - m_expf(-SNOW * DT / MAXSNO)));
      QINTS = fmax2(QINTS, 0.0f);
      const float FT = fmax2(0.0f, (c.TV - 270.15f) / 1.87E5f);
      const float FV = sqrtf(f.UU * f.UU + f.VV * f.VV) / 1.56E5f;
      QDRIPS = fmax2(0.0f, c.CANICE) * (FV + FT);
      QTHROS = (1.0f - fveg) * SNOW + (fveg * SNOW - QINTS);
    } else {
      QINTS = 0.0f; QDRIPS = 0.0f; QTHROS = SNOW;
    }
    QSUBC = fmin2(c.CANICE / DT, QSUBC);                                                  // func:4994-4996
    c.CANICE = fmax2(0.0f, c.CANICE + (QINTS - QDRIPS) * DT + (QFROC - QSUBC) * DT);
    if (c.CANICE <= 1.0E-6f) c.CANICE = 0.0f;
    float FWET;                                                                           // func:5000-5005
    if (c.CANICE > 0.0f) FWET = fmax2(0.0f, c.CANICE) / fmax2(MAXSNO, 1.0E-06f);
    else FWET = fmax2(0.0f, c.CANLIQ) / fmax2(MAXLIQ, 1.0E-06f);
    FWET = m_powf(fmin2(FWET, 1.0f), 0.667f);
    c.FWET = FWET;
    if (c.CANICE > 1.0E-6f && c.TV > TFRZ) {                                              // func:5012-5017
      const float QMELTC = fmin2(c.CANICE / DT, (c.TV - TFRZ) * CICE * c.CANICE / DENICE / (DT * HFUS));
      c.CANICE = fmax2(0.0f, c.CANICE - QMELTC * DT);
      c.CANLIQ = fmax2(0.0f, c.CANLIQ + QMELTC * DT);
      c.TV = FWET * TFRZ + (1.0f - FWET) * c.TV;
    }
    if (c.CANLIQ > 1.0E-6f && c.TV < TFRZ) {                                              // func:5019-5024
      const float QFRZC = fmin2(c.CANLIQ / DT, (TFRZ - c.TV) * CWAT * c.CANLIQ / DENWAT / (DT * HFUS));
      c.CANLIQ = fmax2(0.0f, c.CANLIQ - QFRZC * DT);
      c.CANICE = fmax2(0.0f, c.CANICE + QFRZC * DT);
      c.TV = FWET * TFRZ + (1.0f - FWET) * c.TV;
    }
    w.ECAN = QEVAC + QSUBC - QDEWC - QFROC;                                               // func:5032
    w.ETRAN = ETRAN;
    QRAIN = QDRIPR + QTHROR;                                                              // func:5036-5038
    QSNOW = QDRIPS + QTHROS;
    snowhin = QSNOW / BDFALL;
    if (c.IST == 2 && c.TG > TFRZ) { QSNOW = 0.0f; snowhin = 0.0f; }                      // func:5041-5044
    c.QSNOW = QSNOW;
  }

  NMP_PHASE_SYNC_AT(15);
  // ---- ground evaporation / dew split between snow and soil, func:4725-4735 ----------------------
  float QSNSUB = 0.0f;
  if (c.sneqv > 0.0f) QSNSUB = fmin2(QVAP, c.sneqv / DT);
  float QSEVA = QVAP - QSNSUB;
  float QSNFRO = 0.0f;
  if (c.sneqv > 0.0f) QSNFRO = QDEW;
  float QSDEW = QDEW - QSNFRO;

  // ---- snowwater, func:5049-5174 ------------------------------------------------------------------
  float QSNBOT = 0.0f, PONDING1 = 0.0f, PONDING2 = 0.0f;
  {
    // snowfall, func:5177-5233
    int NEWNODE = 0;
    if (c.ISNOW == 0 && QSNOW > 0.0f) {
      c.snowh = c.snowh + snowhin * DT;
      c.sneqv = c.sneqv + QSNOW * DT;
    }
    if (c.ISNOW == 0 && QSNOW > 0.0f && c.snowh >= 0.025f) {
      c.ISNOW = -1;
      NEWNODE = 1;
      DZSNSO[IX7(0)] = c.snowh;
      c.snowh = 0.0f;
      c.STC[IX7(0)] = fmin2(273.16f, f.SFCTMP);
      c.SNICE[IX3(0)] = c.sneqv;
      c.SNLIQ[IX3(0)] = 0.0f;
    }
    if (c.ISNOW < 0 && NEWNODE == 0 && QSNOW > 0.0f) {
      static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k == c.ISNOW + 1) {
          c.SNICE[IX3(k)] = c.SNICE[IX3(k)] + QSNOW * DT;
          DZSNSO[IX7(k)] = DZSNSO[IX7(k)] + snowhin * DT;
        }
      });
    }
    if (c.ISNOW < 0) {
      // compact, func:5580-5677
      float BURDEN = 0.0f;
      static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k >= c.ISNOW + 1) {
          const float WX = c.SNICE[IX3(k)] + c.SNLIQ[IX3(k)];
          const float FICE = c.SNICE[IX3(k)] / WX;
          const float VOID = 1.0f - (c.SNICE[IX3(k)] / DENICE + c.SNLIQ[IX3(k)] / DENWAT) / DZSNSO[IX7(k)];
          if (VOID > 0.001f && c.SNICE[IX3(k)] > 0.1f) {
            const float BI = c.SNICE[IX3(k)] / DZSNSO[IX7(k)];
            const float TD = fmax2(0.0f, TFRZ - c.STC[IX7(k)]);
            const float DEXPF = m_expf(-0.04f * TD);
            float DDZ1 = -2.5e-6f * DEXPF;
            if (BI > 100.0f) DDZ1 = DDZ1 * m_expf(-46.0E-3f * (BI - 100.0f));
            if (c.SNLIQ[IX3(k)] > 0.01f * DZSNSO[IX7(k)]) DDZ1 = DDZ1 * 2.0f;
            const float DDZ2 = -(BURDEN + 0.5f * WX) * m_expf(-0.08f * TD - 21.e-3f * BI) / 0.8e+6f;
            float DDZ3;
            if (e.IMELT_snow[IX3(k)] == 1) {
              DDZ3 = fmax2(0.0f, (FICEOLD[IX3(k)] - FICE) / fmax2(1.E-6f, FICEOLD[IX3(k)]));
              DDZ3 = -DDZ3 / DT;
            } else {
              DDZ3 = 0.0f;
            }
            float PDZDTC = (DDZ1 + DDZ2 + DDZ3) * DT;
            PDZDTC = fmax2(-0.5f, PDZDTC);
            DZSNSO[IX7(k)] = DZSNSO[IX7(k)] * (1.0f + PDZDTC);
          }
          BURDEN = BURDEN + WX;
        }
      });
      // combine + divide on the scratch copy, func:5109-5117
      SnowPack s;
      pack_snow(c, DZSNSO, s);
      s.ponding1 = PONDING1; s.ponding2 = PONDING2; s.soilice1 = soilice[0];
      combine(s, DZSNSO[IX7(1)]);
      if (s.isnow < 0) divide(s);
      unpack_snow(s, c, DZSNSO);
      PONDING1 = s.ponding1; PONDING2 = s.ponding2; soilice[0] = s.soilice1;
    }
    // snowh2o, func:5680-5819
    if (c.sneqv == 0.0f) {                                                                // func:5726-5732
      soilice[0] = soilice[0] + (QSNFRO - QSNSUB) * DT / (DZSNSO[IX7(1)] * 1000.0f);
      if (soilice[0] < 0.0f) { c.soilwat[0] = c.soilwat[0] + soilice[0]; soilice[0] = 0.0f; }
    }
    if (c.ISNOW == 0 && c.sneqv > 0.0f) {                                                 // func:5739-5754
      const float TEMP = c.sneqv;
      c.sneqv = c.sneqv - QSNSUB * DT + QSNFRO * DT;
      const float PROPOR = c.sneqv / TEMP;
      c.snowh = fmax2(0.0f, PROPOR * c.snowh);
      if (c.sneqv < 0.0f) {
        soilice[0] = soilice[0] + c.sneqv / (DZSNSO[IX7(1)] * 1000.0f);
        c.sneqv = 0.0f;
        c.snowh = 0.0f;
      }
      if (soilice[0] < 0.0f) { c.soilwat[0] = c.soilwat[0] + soilice[0]; soilice[0] = 0.0f; }
    }
    if (c.snowh <= 1.0E-8f || c.sneqv <= 1.0E-6f) { c.snowh = 0.0f; c.sneqv = 0.0f; }     // func:5756-5759
    if (c.ISNOW < 0) {                                                                    // func:5763-5778
      float WGDIF = 0.0f;
      static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k == c.ISNOW + 1) {
          WGDIF = c.SNICE[IX3(k)] - QSNSUB * DT + QSNFRO * DT;
          c.SNICE[IX3(k)] = WGDIF;
        }
      });
      if (WGDIF < 1.0E-6f) {
        SnowPack s;
        pack_snow(c, DZSNSO, s);
        s.ponding1 = PONDING1; s.ponding2 = PONDING2; s.soilice1 = soilice[0];
        combine(s, DZSNSO[IX7(1)]);
        unpack_snow(s, c, DZSNSO);
        PONDING1 = s.ponding1; PONDING2 = s.ponding2; soilice[0] = s.soilice1;
      }
      if (c.ISNOW < 0) {
        static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
          if (k == c.ISNOW + 1) {
            c.SNLIQ[IX3(k)] = c.SNLIQ[IX3(k)] + QRAIN * DT;
            c.SNLIQ[IX3(k)] = fmax2(0.0f, c.SNLIQ[IX3(k)]);
          }
        });
      }
    }
    {                                                                                     // func:5784-5818 percolation
      float VOL_LIQ[3] = {0.0f, 0.0f, 0.0f}, VOL_ICE[3] = {0.0f, 0.0f, 0.0f}, epore[3] = {0.0f, 0.0f, 0.0f};
      static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k >= c.ISNOW + 1) {
          VOL_ICE[IX3(k)] = fmin2(1.0f, c.SNICE[IX3(k)] / (DZSNSO[IX7(k)] * DENICE));
          epore[IX3(k)] = 1.0f - VOL_ICE[IX3(k)];
          VOL_LIQ[IX3(k)] = fmin2(epore[IX3(k)], c.SNLIQ[IX3(k)] / (DZSNSO[IX7(k)] * DENWAT));
        }
      });
      float QIN = 0.0f, QOUT = 0.0f;
      static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k >= c.ISNOW + 1) {
          c.SNLIQ[IX3(k)] = c.SNLIQ[IX3(k)] + QIN;
          if (k <= -1) {
            if (epore[IX3(k)] < 0.05f || epore[IX3(k < 0 ? k + 1 : k)] < 0.05f) {
              QOUT = 0.0f;
            } else {
              QOUT = fmax2(0.0f, (VOL_LIQ[IX3(k)] - T.gen[G_SSI] * epore[IX3(k)]) * DZSNSO[IX7(k)]);
              QOUT = fmin2(QOUT, (1.0f - VOL_ICE[IX3(k < 0 ? k + 1 : k)] - VOL_LIQ[IX3(k < 0 ? k + 1 : k)]) *
                                     DZSNSO[IX7(k + 1)]);
            }
          } else {
            QOUT = fmax2(0.0f, (VOL_LIQ[IX3(k)] - T.gen[G_SSI] * epore[IX3(k)]) * DZSNSO[IX7(k)]);
          }
          QOUT = QOUT * 1000.0f;
          c.SNLIQ[IX3(k)] = c.SNLIQ[IX3(k)] - QOUT;
          QIN = QOUT;
        }
      });
      QSNBOT = QOUT / DT;
    }
    // back in snowwater: func:5127-5172
    static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
      if (k <= c.ISNOW) {
        c.SNICE[IX3(k)] = 0.0f; c.SNLIQ[IX3(k)] = 0.0f; c.STC[IX7(k)] = 0.0f; DZSNSO[IX7(k)] = 0.0f; c.ZSNSO[IX7(k)] = 0.0f;
      }
    });
    if (c.sneqv > 2000.0f) {                                                              // func:5137-5143
      const float BDSNOW = c.SNICE[IX3(0)] / DZSNSO[IX7(0)];
      SNOFLOW = (c.sneqv - 2000.0f);
      c.SNICE[IX3(0)] = c.SNICE[IX3(0)] - SNOFLOW;
      DZSNSO[IX7(0)] = DZSNSO[IX7(0)] - SNOFLOW / BDSNOW;
      SNOFLOW = SNOFLOW / DT;
    }
    if (c.ISNOW < 0) {                                                                    // func:5147-5152
      c.sneqv = 0.0f;
      static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (k >= c.ISNOW + 1) c.sneqv = c.sneqv + c.SNICE[IX3(k)] + c.SNLIQ[IX3(k)];
      });
    }
                                                       static_for<-2, 0>([&](auto k_) { constexpr int k = decltype(k_)::value;
                                                         // func:5156-5158
      if (k >= c.ISNOW + 1) DZSNSO[IX7(k)] = -DZSNSO[IX7(k)];
    });
    DZSNSO[IX7(1)] = rc.zsoil[0];                                                         // func:5160-5163
    static_for<2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      DZSNSO[IX7(k)] = (rc.zsoil[k - 1] - rc.zsoil[k - 2]);
    });
    static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;                                                   // func:5165-5168
      if (k == c.ISNOW + 1) c.ZSNSO[IX7(k)] = DZSNSO[IX7(k)];
      else if (k > c.ISNOW + 1) c.ZSNSO[IX7(k)] = c.ZSNSO[IX7(k > -2 ? k - 1 : k)] + DZSNSO[IX7(k)];
    });
                                                   static_for<-2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
                                                     // func:5170-5172
      if (k >= c.ISNOW + 1) DZSNSO[IX7(k)] = -DZSNSO[IX7(k)];
    });
  }
  w.QSNBOT = QSNBOT; w.PONDING1 = PONDING1; w.PONDING2 = PONDING2;

  NMP_PHASE_SYNC_AT(16);
  if (e.frozen_ground) {                                                                  // func:4744-4752
    soilice[0] = soilice[0] + (QSDEW - QSEVA) * DT / (DZSNSO[IX7(1)] * 1000.0f);
    QSDEW = 0.0f;
    QSEVA = 0.0f;
    if (soilice[0] < 0.0f) { c.soilwat[0] = c.soilwat[0] + soilice[0]; soilice[0] = 0.0f; }
  }
  float qinsrf = (e.PONDING + PONDING1 + PONDING2) / DT * 0.001f;                         // func:4757-4764
  if (c.ISNOW == 0) qinsrf = qinsrf + (QSNBOT + QSDEW + QRAIN) * 0.001f;
  else qinsrf = qinsrf + (QSNBOT + QSDEW) * 0.001f;
  QSEVA = QSEVA * 0.001f;                                                                 // func:4766
  const int NROOT = T.vegi[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGI + VI_NROOT];
  float ETRANI[4];
  static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
    ETRANI[IX4(k)] = (k <= NROOT) ? w.ETRAN * e.BTRANI[IX4(k)] * 0.001f : 0.0f;
  });  // func:4768-4770

  if (c.IST == 2) {                                                                       // func:4774-4777 lake bucket
    w.runsrf = 0.0f;
    if (c.WSLAKE >= 5000.0f) w.runsrf = qinsrf * 1000.0f;
    c.WSLAKE = c.WSLAKE + (qinsrf - QSEVA) * 1000.0f * DT - w.runsrf * DT;
  }
  // ---- soilh2o, func:5822-6048 (soil columns; barriers only at CTA-convergent points) -----------------
  const bool soil_col = (c.IST != 2);
  float FCR[4] = {0.0f, 0.0f, 0.0f, 0.0f}, WCND[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  float runsrf = 0.0f, qinfil = 0.0f, RSAT = 0.0f, QDRAIN = 0.0f, runsub = 0.0f;
  float SICEMAX = 0.0f, FCRMAX = 0.0f;
  int niter = 1;
  float dtfine = DT, QDRAIN_SAVE = 0.0f;
  NMP_PHASE_SYNC_AT(17);
  if (soil_col) {
    static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;                                                    // func:5893-5897
      const float epore = fmax2(1.0E-4f, (SMCMAX - soilice[IX4(k)]));
      RSAT = RSAT + fmax2(0.0f, c.soilwat[IX4(k)] - epore) * DZSNSO[IX7(k)];
      c.soilwat[IX4(k)] = fmin2(epore, c.soilwat[IX4(k)]);
    });
    {
      const float em4 = m_expf(-4.0f);
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;                                                  // func:5900-5904
        const float FICE = fmin2(1.0f, soilice[IX4(k)] / SMCMAX);
        FCR[IX4(k)] = fmax2(0.0f, m_expf(-4.0f * (1.0f - FICE)) - em4) / (1.0f - em4);
      });
    }
    static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;                                                    // func:5907-5914
      if (soilice[IX4(k)] > SICEMAX) SICEMAX = soilice[IX4(k)];
      if (FCR[IX4(k)] > FCRMAX) FCRMAX = FCR[IX4(k)];
    });
    if (opt.run == 2) {                                                                   // func:5917-5922 + zwteq 6051-6100
      float WD1 = 0.0f;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        WD1 = WD1 + (SMCMAX - c.soilwat[IX4(k)]) * DZSNSO[IX7(k)];
      });
      const float DZFINE = 3.0f * (-rc.zsoil[3]) / 100.0f;
      float ZWT = -3.0f * rc.zsoil[3] - 0.001f;
      float WD2 = 0.0f;
#pragma unroll 1
      for (int K = 1; K <= 100; ++K) {
        const float ZF = (float)K * DZFINE;
        const float TEMP = 1.0f + (ZWT - ZF) / SP[P_PSISAT];
        WD2 = WD2 + SMCMAX * (1.0f - m_powf(TEMP, -1.0f / BEXP)) * DZFINE;
        if (fabsf(WD2 - WD1) <= 0.01f) { ZWT = ZF; break; }
      }
      c.ZWT = ZWT;
      runsub = (1.0f - FCRMAX) * 4.0f * m_expf(-T.gen[G_TIMEAN]) * m_expf(-2.0f * c.ZWT);
    }
    if (c.lutyp == ISURBAN) FCR[0] = 0.95f;                                               // func:5927
    if (opt.run == 1) {                                                                   // func:5929-5936
      const float FSAT = T.gen[G_FSATMAX] * m_expf(-0.5f * 6.0f * (c.ZWT - 2.0f));
      if (qinsrf > 0.0f) { runsrf = qinsrf * ((1.0f - FCR[0]) * FSAT + FCR[0]); qinfil = qinsrf - runsrf; }
    }
    if (opt.run == 2) {                                                                   // func:5938-5945
      const float FSAT = T.gen[G_FSATMAX] * m_expf(-0.5f * 2.0f * c.ZWT);
      if (qinsrf > 0.0f) { runsrf = qinsrf * ((1.0f - FCR[0]) * FSAT + FCR[0]); qinfil = qinsrf - runsrf; }
    }
    if (opt.run == 3) {                                                                   // infil, func:6103-6196
      if (qinsrf > 0.0f) {
        const float DT1 = DT / 86400.0f;
        const float SMCAV = SMCMAX - SP[P_SMCWLT];
        float DD = 0.0f, DICE = 0.0f;
        static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
          const float thick = (k == 1) ? -rc.zsoil[0] : (rc.zsoil[k > 1 ? k - 2 : 0] - rc.zsoil[k - 1]);
          float DMAX = thick * SMCAV;
          const float dice_k = thick * soilice[IX4(k)];
          DICE = (k == 1) ? dice_k : DICE + dice_k;
          DMAX = DMAX * (1.0f - (c.soilwat[IX4(k)] + soilice[IX4(k)] - SP[P_SMCWLT]) / SMCAV);
          DD = (k == 1) ? DMAX : DD + DMAX;
        });
        const float VAL = (1.0f - m_expf(-SP[P_KDT] * DT1));
        const float DDT = DD * VAL;
        const float PX = fmax2(0.0f, qinsrf * DT);
        float INFMAX = (PX * (DDT / (PX + DDT))) / DT;
        float FCRi = 1.0f;
        if (DICE > 1.0E-2f) {
          const float ACRT = 3.0f * SP[P_FRZX] / DICE;
          float SUM = 1.0f;
          SUM = SUM + (ACRT * ACRT) / 2.0f;   // J=1: K = 2
          SUM = SUM + ACRT / 1.0f;            // J=2: K = 1
          FCRi = 1.0f - m_expf(-ACRT) * SUM;
        }
        INFMAX = INFMAX * FCRi;
        float WDFt, WCNDt;
        wdfcnd2(SMCMAX, BEXP, SP[P_DWSAT], SP[P_DKSAT], c.soilwat[0], SICEMAX, WDFt, WCNDt);
        INFMAX = fmax2(INFMAX, WCNDt);
        INFMAX = fmin2(INFMAX, PX);
        runsrf = fmax2(0.0f, qinsrf - INFMAX);
        qinfil = qinsrf - runsrf;
      }
    }
    if (opt.run == 4) {                                                                   // func:5953-5968
      float SMCTOT = 0.0f, DZTOT = 0.0f;
      bool done = false;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (!done) {
          DZTOT = DZTOT + DZSNSO[IX7(k)];
          SMCTOT = SMCTOT + c.SMC[IX4(k)] * DZSNSO[IX7(k)];
          if (DZTOT >= 2.0f) done = true;
        }
      });
      SMCTOT = SMCTOT / DZTOT;
      const float FSAT = m_powf(fmax2(0.01f, SMCTOT / SMCMAX), 4.0f);
      if (qinsrf > 0.0f) { runsrf = qinsrf * ((1.0f - FCR[0]) * FSAT + FCR[0]); qinfil = qinsrf - runsrf; }
    }
    niter = 1;                                                                            // func:5971-5978
    if (opt.inf == 1) {
      niter = 3;
      if (qinfil * DT > DZSNSO[IX7(1)] * SMCMAX) niter = niter * 2;
    }
    dtfine = DT / (float)niter;
    w.niter_soil = niter;
  }
#pragma unroll 1
  for (int iter = 1; iter <= 6; ++iter) {                                                 // func:5982-5996
    if (!NMP_ANY(soil_col && iter <= niter)) break;
    if (soil_col && iter <= niter) {
      // srt, func:6199-6305
      float WDF[4], SMX[4], AI[4], BI[4], CI[4], RHS[4];
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (opt.inf == 1) {
          wdfcnd1(SMCMAX, BEXP, SP[P_DWSAT], SP[P_DKSAT], c.SMC[IX4(k)], FCR[IX4(k)], WDF[IX4(k)], WCND[IX4(k)]);
          SMX[IX4(k)] = c.SMC[IX4(k)];
        } else {
          wdfcnd2(SMCMAX, BEXP, SP[P_DWSAT], SP[P_DKSAT], c.soilwat[IX4(k)], SICEMAX, WDF[IX4(k)], WCND[IX4(k)]);
          SMX[IX4(k)] = c.soilwat[IX4(k)];
        }
      });
      float dsmdz_prev = 0.0f, ddz_prev = 0.0f;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        float DENOM, DDZ = 0.0f, DSMDZ = 0.0f, WFLUX;
        if (k == 1) {
          DENOM = -rc.zsoil[0];
          const float TEMP1 = -rc.zsoil[1];
          DDZ = 2.0f / TEMP1;
          DSMDZ = 2.0f * (SMX[0] - SMX[1]) / TEMP1;
          WFLUX = WDF[0] * DSMDZ + WCND[0] - qinfil + ETRANI[0] + QSEVA;
          AI[0] = 0.0f;
          BI[0] = WDF[0] * DDZ / DENOM;
          CI[0] = -BI[0];
        } else if (k < NSOIL) {
          DENOM = (rc.zsoil[k - 2 >= 0 ? k - 2 : 0] - rc.zsoil[k - 1]);
          const float TEMP1 = (rc.zsoil[k - 2 >= 0 ? k - 2 : 0] - rc.zsoil[k < NSOIL ? k : k - 1]);
          DDZ = 2.0f / TEMP1;
          DSMDZ = 2.0f * (SMX[IX4(k)] - SMX[IX4(k < NSOIL ? k + 1 : k)]) / TEMP1;
          WFLUX = WDF[IX4(k)] * DSMDZ + WCND[IX4(k)] - WDF[IX4(k > 1 ? k - 1 : k)] * dsmdz_prev -
                  WCND[IX4(k > 1 ? k - 1 : k)] + ETRANI[IX4(k)];
          AI[IX4(k)] = -WDF[IX4(k > 1 ? k - 1 : k)] * ddz_prev / DENOM;
          CI[IX4(k)] = -WDF[IX4(k)] * DDZ / DENOM;
          BI[IX4(k)] = -(AI[IX4(k)] + CI[IX4(k)]);
        } else {
          DENOM = (rc.zsoil[k - 2 >= 0 ? k - 2 : 0] - rc.zsoil[k - 1]);
          if (opt.run == 1 || opt.run == 2) QDRAIN = 0.0f;
          if (opt.run == 3) QDRAIN = T.gen[G_SLOPE + NMP_CHK(c.slptyp - 1, 0, NGEN - G_SLOPE - 1, 4)] * WCND[IX4(k)];
          if (opt.run == 4) QDRAIN = (1.0f - FCRMAX) * WCND[IX4(k)];
          WFLUX = -(WDF[IX4(k > 1 ? k - 1 : k)] * dsmdz_prev) - WCND[IX4(k > 1 ? k - 1 : k)] + ETRANI[IX4(k)] + QDRAIN;
          AI[IX4(k)] = -WDF[IX4(k > 1 ? k - 1 : k)] * ddz_prev / DENOM;
          CI[IX4(k)] = 0.0f;
          BI[IX4(k)] = -(AI[IX4(k)] + CI[IX4(k)]);
        }
        RHS[IX4(k)] = WFLUX / (-DENOM);
        dsmdz_prev = DSMDZ;
        ddz_prev = DDZ;
      });
      // sstep (func:6308-6383) + rosr12 (func:4240-4288) on 4 unknowns
      float P[4], DELTA[4];
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        const float rhs = RHS[IX4(k)] * dtfine, a = AI[IX4(k)] * dtfine, b = 1.0f + BI[IX4(k)] * dtfine;
        const float cc = (k == NSOIL) ? 0.0f : CI[IX4(k)] * dtfine;
        if (k == 1) {
          P[0] = -cc / b;
          DELTA[0] = rhs / b;
        } else {
          P[IX4(k)] = -cc * (1.0f / (b + a * P[IX4(k > 1 ? k - 1 : k)]));
          DELTA[IX4(k)] = (rhs - a * DELTA[IX4(k > 1 ? k - 1 : k)]) * (1.0f / (b + a * P[IX4(k > 1 ? k - 1 : k)]));
        }
      });
      P[3] = DELTA[3];
      static_for_rev<NSOIL - 1, 1>([&](auto k_) { constexpr int k = decltype(k_)::value;
        P[IX4(k)] = P[IX4(k)] * P[IX4(k + 1)] + DELTA[IX4(k)];
      });
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        c.soilwat[IX4(k)] = c.soilwat[IX4(k)] + P[IX4(k)];
      });
      float WPLUS = 0.0f;
      static_for_rev<NSOIL, 2>([&](auto k_) { constexpr int k = decltype(k_)::value;                                                  // func:6372-6377
        const float epore = fmax2(1.0E-4f, SMCMAX - soilice[IX4(k)]);
        WPLUS = fmax2(c.soilwat[IX4(k)] - epore, 0.0f) * DZSNSO[IX7(k)];
        c.soilwat[IX4(k)] = fmin2(epore, c.soilwat[IX4(k)]);
        c.soilwat[IX4(k - 1)] = c.soilwat[IX4(k - 1)] + WPLUS / DZSNSO[IX7(k - 1)];
      });
      {
        const float epore = fmax2(1.0E-4f, SMCMAX - soilice[0]);                          // func:6379-6382
        WPLUS = fmax2(c.soilwat[0] - epore, 0.0f) * DZSNSO[IX7(1)];
        c.soilwat[0] = fmin2(epore, c.soilwat[0]);
      }
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        c.SMC[IX4(k)] = c.soilwat[IX4(k)] + soilice[IX4(k)];
      });
      RSAT = RSAT + WPLUS;
      QDRAIN_SAVE = QDRAIN_SAVE + QDRAIN;
    }
  }
  NMP_PHASE_SYNC_AT(18);
  if (soil_col) {
    QDRAIN = QDRAIN_SAVE / (float)niter;                                                  // func:5998-6001
    runsrf = runsrf * 1000.0f + RSAT * 1000.0f / DT;
    QDRAIN = QDRAIN * 1000.0f;
    if (opt.run == 2) {                                                                   // func:6004-6014
      float WTSUB = 0.0f;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        WTSUB = WTSUB + WCND[IX4(k)] * DZSNSO[IX7(k)];
      });
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        const float MH2O = runsub * DT * (WCND[IX4(k)] * DZSNSO[IX7(k)]) / WTSUB;
        c.soilwat[IX4(k)] = c.soilwat[IX4(k)] - MH2O / (DZSNSO[IX7(k)] * 1000.0f);
      });
    }
    if (opt.run != 1) {                                                                   // func:6018-6046
      float MLIQ[4];
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        MLIQ[IX4(k)] = c.soilwat[IX4(k)] * DZSNSO[IX7(k)] * 1000.0f;
      });
      const float WATMIN = 0.01f;
      static_for<1, NSOIL - 1>([&](auto k_) { constexpr int k = decltype(k_)::value;
        const float XS = (MLIQ[IX4(k)] < 0.0f) ? WATMIN - MLIQ[IX4(k)] : 0.0f;
        MLIQ[IX4(k)] = MLIQ[IX4(k)] + XS;
        MLIQ[IX4(k + 1)] = MLIQ[IX4(k + 1)] - XS;
      });
      const float XS = (MLIQ[3] < WATMIN) ? WATMIN - MLIQ[3] : 0.0f;
      MLIQ[3] = MLIQ[3] + XS;
      runsub = runsub - XS / DT;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        c.soilwat[IX4(k)] = div_rn(MLIQ[IX4(k)], DZSNSO[IX7(k)] * 1000.0f);
      });
    }
    w.runsrf = runsrf;
    w.runsub = runsub;

    // ---- groundwater (SIMGM), func:6458-6639 ------------------------------------------------------
    if (opt.run == 1) {
      float DZMM[4], ZNODE[4], MLIQ[4], epore[4], HK[4], SMCg[4];
      DZMM[0] = -rc.zsoil[0] * 1.0E3f;
      ZNODE[0] = -rc.zsoil[0] / 2.0f;
      static_for<2, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        DZMM[IX4(k)] = 1.0E3f * (rc.zsoil[k - 2] - rc.zsoil[k - 1]);
        ZNODE[IX4(k)] = -rc.zsoil[k - 2] + 0.5f * (rc.zsoil[k - 2] - rc.zsoil[k - 1]);
      });
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        SMCg[IX4(k)] = c.soilwat[IX4(k)] + soilice[IX4(k)];
        MLIQ[IX4(k)] = c.soilwat[IX4(k)] * DZMM[IX4(k)];
        epore[IX4(k)] = fmax2(0.01f, SMCMAX - soilice[IX4(k)]);
        HK[IX4(k)] = 1.0E3f * WCND[IX4(k)];
      });
      int IWT = NSOIL;                                                                    // func:6545-6551
      static_for_rev<NSOIL, 2>([&](auto k_) { constexpr int k = decltype(k_)::value;
        if (c.ZWT <= -rc.zsoil[k - 1]) IWT = k - 1;
      });
      const float QDIS = (1.0f - FCRMAX) * 5.0f * m_expf(-T.gen[G_TIMEAN]) * m_expf(-6.0f * (c.ZWT - 2.0f));  // func:6557
#ifdef NMP_FAST_ALGEBRA   // the fast build keeps this power in fp32, but with libdevice's accurate routine (not the MUFU pair)
      const float S_NODEf = fmax2(fmin2(1.0f, pick(SMCg, IWT - 1, 6) / SMCMAX), 0.01f);
      float SMPFZ = -((SP[P_PSISAT] * 1000.0f) * m_accpowf(S_NODEf, -BEXP));
#else
      double S_NODE = (double)fmin2(1.0f, pick(SMCg, IWT - 1, 6) / SMCMAX);                  // func:6560-6563 (fp64 island)
      const double S_MIN = (double)0.01f;
      S_NODE = S_NODE > S_MIN ? S_NODE : S_MIN;
      float SMPFZ = (float)(-((double)(SP[P_PSISAT] * 1000.0f) * pow(S_NODE, (double)(-BEXP))));
#endif
      SMPFZ = fmax2(-120000.0f, 0.20f * SMPFZ);
      const float KA = pick(HK, IWT - 1, 6);
      const float znode_iwt = pick(ZNODE, IWT - 1, 6);
      const float WH_ZWT = -c.ZWT * 1.0E3f;                                               // func:6568-6571
      const float WH = SMPFZ - znode_iwt * 1.0E3f;
      float QIN = -KA * (WH_ZWT - WH) / ((c.ZWT - znode_iwt) * 1.0E3f);
      QIN = fmax2(-10.0f / DT, fmin2(10.0f / DT, QIN));
      c.WT = c.WT + (QIN - QDIS) * DT;                                                    // func:6575
      if (IWT == NSOIL) {                                                                 // func:6577-6584
        c.WA = c.WA + (QIN - QDIS) * DT;
        c.WT = c.WA;
        c.ZWT = (-rc.zsoil[3] + 25.0f) - c.WA / 1000.0f / 0.2f;
        MLIQ[3] = MLIQ[3] - QIN * DT;
        MLIQ[3] = MLIQ[3] + fmax2(0.0f, c.WA - 5000.0f);
        c.WA = fmin2(c.WA, 5000.0f);
      } else {
        if (IWT == NSOIL - 1) {                                                           // func:6587-6597
          c.ZWT = -rc.zsoil[3] - (c.WT - 0.2f * 1000.0f * 25.0f) / (epore[3]) / 1000.0f;
        } else {
          float WS = 0.0f;
          static_for<3, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
            if (k >= IWT + 2) WS = WS + epore[IX4(k)] * DZMM[IX4(k)];
          });
          c.ZWT = -pick(rc.zsoil, IWT, 6) - (c.WT - 0.2f * 1000.0f * 25.0f - WS) / (pick(epore, IWT, 6)) / 1000.0f;
        }
        float WTSUB = 0.0f;                                                               // func:6599-6606
        static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
          WTSUB = WTSUB + HK[IX4(k)] * DZMM[IX4(k)];
        });
        static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
          MLIQ[IX4(k)] = MLIQ[IX4(k)] - QDIS * DT * HK[IX4(k)] * DZMM[IX4(k)] / WTSUB;
        });
      }
      c.ZWT = fmax2(1.5f, c.ZWT);                                                         // func:6609
      const float WATMIN = 0.01f;                                                         // func:6615-6638
      static_for<1, NSOIL - 1>([&](auto k_) { constexpr int k = decltype(k_)::value;
        const float XS = (MLIQ[IX4(k)] < 0.0f) ? WATMIN - MLIQ[IX4(k)] : 0.0f;
        MLIQ[IX4(k)] = MLIQ[IX4(k)] + XS;
        MLIQ[IX4(k + 1)] = MLIQ[IX4(k + 1)] - XS;
      });
      const float XS = (MLIQ[3] < WATMIN) ? WATMIN - MLIQ[3] : 0.0f;
      MLIQ[3] = MLIQ[3] + XS;
      c.WA = c.WA - XS;
      c.WT = c.WT - XS;
      static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
        c.soilwat[IX4(k)] = div_rn(MLIQ[IX4(k)], DZMM[IX4(k)]);
      });
      w.runsub = QDIS;                                                                    // func:4790
    }
    if (opt.run == 3 || opt.run == 4) w.runsub = w.runsub + QDRAIN;                       // func:4793-4795
    static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
      c.SMC[IX4(k)] = c.soilwat[IX4(k)] + soilice[IX4(k)];
    }); // func:4797-4799
  }
  w.runsub = w.runsub + SNOFLOW;                                                          // func:4802
}

// =====================================================================================================
// carbon / co2flux, func:6642-7025 (dynamic vegetation, opt_veg 2 or 5 only)
// =====================================================================================================
__device__ __forceinline__ void carbon(const Tables& T, const RunConst& rc, Column& c, const float (&DZSNSO)[7],
                                    float PSN, float BTRAN, float IGS, float& GPP, float& NPP, float& NEE) {
  const float* VG = &T.veg[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGF];
  const float DT = rc.dt;
  if ((c.lutyp == T.flags[1]) || (c.lutyp == T.flags[2]) || (c.lutyp == T.flags[3]) || (c.lutyp == T.flags[0])) {
    c.LAI = 0.0f; c.SAI = 0.0f; GPP = 0.0f; NPP = 0.0f; NEE = 0.0f;                       // func:6709-6726
    c.LFMASS = 0.0f; c.RTMASS = 0.0f; c.STMASS = 0.0f; c.WOOD = 0.0f; c.STBLCP = 0.0f; c.FASTCP = 0.0f;
    return;
  }
  const float SMCMAX = T.soil[NMP_CHK(c.sltyp - 1, 0, MAXSOILTYP - 1, 4) * NSOILP + P_SMCMAX];
  const float LAPM = VG[V_SLA] / 1000.0f;
  const float WSTRES = 1.0f - BTRAN;
  const int NROOT = T.vegi[NMP_CHK(c.lutyp - 1, 0, MAXVEGTYP - 1, 4) * NVEGI + VI_NROOT];
  float WROOT = 0.0f;
  const float zroot = root_depth(rc, NROOT);
  static_for<1, NSOIL>([&](auto k_) { constexpr int k = decltype(k_)::value;
    if (k <= NROOT) WROOT = WROOT + c.SMC[IX4(k)] / SMCMAX * DZSNSO[IX7(k)] / zroot;
  });
  // co2flux, func:6754-7025
  const float RTOVRC = 2.0E-8f, RSWOODC = 3.0E-10f, BF = 0.90f, WSTRC = 100.0f, LAIMIN = 0.05f, XSAMIN = 0.01f;
  const float SAPM = 3.0f * 0.001f;
  const float LFMSMN = LAIMIN / LAPM, STMSMN = XSAMIN / SAPM;
  const float RF = (IGS == 0.0f) ? 0.5f : 1.0f;
  const float FNF = fmin2(rc.foln / fmax2(1.0E-06f, VG[V_FOLNMX]), 1.0f);
  const float TF = m_powf(VG[V_ARM], (c.TV - 298.16f) / 10.0f);
  const float RESP = VG[V_RMF25] * TF * FNF * c.LAI * RF * (1.0f - WSTRES);
  const float RSLEAF = fmin2(c.LFMASS / DT, RESP * 12.0E-6f);
  const float RSROOT = VG[V_RMR25] * (c.RTMASS * 1.0E-3f) * TF * RF * 12.0E-6f;
  const float RSSTEM = VG[V_RMS25] * (c.STMASS * 1.0E-3f) * TF * RF * 12.0E-6f;
  const float RSWOOD = RSWOODC * m_expf(0.08f * (c.TV - 298.16f)) * c.WOOD * VG[V_WDPOOL];
  const float CARBFX = PSN * 12.0E-6f;
  float LEAFPT = m_expf(0.01f * (1.0f - m_expf(0.75f * c.LAI)) * c.LAI);
  if (c.lutyp == T.flags[4]) LEAFPT = m_expf(0.01f * (1.0f - m_expf(0.50f * c.LAI)) * c.LAI);
  const float NONLEF = 1.0f - LEAFPT;
  const float STEMPT = c.LAI / 10.0f;
  LEAFPT = LEAFPT - STEMPT;
  float WOODF;
  if (c.WOOD > 0.0f) WOODF = (1.0f - m_expf(-BF * (VG[V_WRRAT] * c.RTMASS / c.WOOD)) / BF) * VG[V_WDPOOL];
  else WOODF = 0.0f;
  const float ROOTPT = NONLEF * (1.0f - WOODF), WOODPT = NONLEF * WOODF;
  const float LFTOVR = VG[V_LTOVRC] * 1.0E-6f * c.LFMASS, STTOVR = VG[V_LTOVRC] * 1.0E-6f * c.STMASS;
  float RTTOVR = RTOVRC * c.RTMASS;
  const float WDTOVR = 9.5E-10f * c.WOOD;
  const float SC = m_expf(-0.3f * fmax2(0.0f, c.TV - VG[V_TDLEF])) * (c.LFMASS / 120.0f);
  const float SD = m_expf((WSTRES - 1.0f) * WSTRC);
  float DIELF = c.LFMASS * 1.0E-6f * (VG[V_DILEFW] * SD + VG[V_DILEFC] * SC);
  float DIEST = c.STMASS * 1.0E-6f * (VG[V_DILEFW] * SD + VG[V_DILEFC] * SC);
  const float GRLEAF = fmax2(0.0f, VG[V_FRAGR] * (LEAFPT * CARBFX - RSLEAF));
  const float GRSTEM = fmax2(0.0f, VG[V_FRAGR] * (STEMPT * CARBFX - RSSTEM));
  const float GRROOT = fmax2(0.0f, VG[V_FRAGR] * (ROOTPT * CARBFX - RSROOT));
  const float GRWOOD = fmax2(0.0f, VG[V_FRAGR] * (WOODPT * CARBFX - RSWOOD));
  float ADDNPPLF = fmax2(0.0f, LEAFPT * CARBFX - GRLEAF - RSLEAF);
  float ADDNPPST = fmax2(0.0f, STEMPT * CARBFX - GRSTEM - RSSTEM);
  if (c.TV < VG[V_TMIN]) ADDNPPLF = 0.0f;
  if (c.TV < VG[V_TMIN]) ADDNPPST = 0.0f;
  const float LFDEL = (c.LFMASS - LFMSMN) / DT, STDEL = (c.STMASS - STMSMN) / DT;
  DIELF = fmin2(DIELF, LFDEL + ADDNPPLF - LFTOVR);
  DIEST = fmin2(DIEST, STDEL + ADDNPPST - STTOVR);
  const float NPPL = fmax2(ADDNPPLF, -LFDEL), NPPS = fmax2(ADDNPPST, -STDEL);
  const float NPPR = ROOTPT * CARBFX - RSROOT - GRROOT, NPPW = WOODPT * CARBFX - RSWOOD - GRWOOD;
  c.LFMASS = c.LFMASS + (NPPL - LFTOVR - DIELF) * DT;
  c.STMASS = c.STMASS + (NPPS - STTOVR - DIEST) * DT;
  c.RTMASS = c.RTMASS + (NPPR - RTTOVR) * DT;
  if (c.RTMASS < 0.0f) { RTTOVR = NPPR; c.RTMASS = 0.0f; }
  c.WOOD = (c.WOOD + (NPPW - WDTOVR) * DT) * VG[V_WDPOOL];
  c.FASTCP = c.FASTCP + (RTTOVR + LFTOVR + STTOVR + WDTOVR + DIELF) * DT;
  const float FST = m_powf(2.0f, (c.STC[IX7(1)] - 283.16f) / 10.0f);
  const float FSW = WROOT / (0.20f + WROOT) * 0.23f / (0.23f + WROOT);
  const float RSSOIL = FSW * FST * VG[V_MRP] * fmax2(0.0f, c.FASTCP * 1.0E-3f) * 12.0E-6f;
  const float STABLC = 0.1f * RSSOIL;
  c.FASTCP = c.FASTCP - (RSSOIL + STABLC) * DT;
  c.STBLCP = c.STBLCP + STABLC * DT;
  GPP = CARBFX;
  NPP = NPPL + NPPW + NPPR;
  const float AUTORS = RSROOT + RSWOOD + RSLEAF + GRLEAF + GRROOT + GRWOOD;
  NEE = (AUTORS + RSSOIL - GPP) * 44.0f / 12.0f;
  c.LAI = fmax2(c.LFMASS * LAPM, LAIMIN);
  c.SAI = fmax2(c.STMASS * SAPM, XSAMIN);
}

}  // namespace nmp
