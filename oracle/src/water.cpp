// TEST INFRASTRUCTURE ONLY (see common.h).
// water, canwater, snowwater, snowfall, combine, divide, combo, compact, snowh2o :
// core/module_noahmp_func.f90:4601-5819
#include "physics.h"

namespace noahmp_oracle {

// func:4601-4804
void water(const Ctx& cx, ColumnIO& c, const I7& IMELT, float QPRECC, float QPRECL, float elai,
           float esai, float QVAP, float QDEW, const A4& BTRANI, float PONDING, bool frozen_canopy,
           bool frozen_ground, A4& soilice, A7& DZSNSO, float& CMC, float& QIN, float& QDIS) {
  const Tables& t = cx.tbl;
  const float DT = c.DT;
  const float WSLMAX = 5000.0f;                                 // func:4705
  A4 ETRANI, WCND;
  for (int IZ = 1; IZ <= NSOIL; ++IZ) { ETRANI(IZ) = 0.0f; WCND(IZ) = 0.0f; }  // func:4709
  float SNOFLOW = 0.0f;
  c.runsub = 0.0f;
  float qinsrf = 0.0f;
  QIN = 0.0f;
  QDIS = 0.0f;

  float QRAIN, snowhin;
  canwater(cx, c.lutyp, DT, c.SFCTMP, c.UU, c.VV, c.FCEV, c.FCTR, QPRECC, QPRECL, elai, esai, c.IST,
           c.TG, c.fveg, frozen_canopy, c.CANLIQ, c.CANICE, c.TV, CMC, c.ECAN, c.ETRAN, QRAIN, c.QSNOW,
           snowhin, c.FWET, c.FPICE);                           // func:4715-4721

  float QSNSUB = 0.0f;                                          // func:4725-4735
  if (c.sneqv > 0.0f) QSNSUB = fmin2(QVAP, c.sneqv / DT);
  float QSEVA = QVAP - QSNSUB;
  float QSNFRO = 0.0f;
  if (c.sneqv > 0.0f) QSNFRO = QDEW;
  float QSDEW = QDEW - QSNFRO;

  snowwater(cx, DT, c.ZSOIL, IMELT, c.SFCTMP, snowhin, c.QSNOW, QSNFRO, QSNSUB, QRAIN, c.FICEOLD,
            c.ISNOW, c.snowh, c.sneqv, c.SNICE, c.SNLIQ, c.soilwat, soilice, c.STC, c.ZSNSO, DZSNSO,
            c.QSNBOT, SNOFLOW, c.PONDING1, c.PONDING2);         // func:4737-4742

  if (frozen_ground) {                                          // func:4744-4752
    soilice(1) = soilice(1) + (QSDEW - QSEVA) * DT / (DZSNSO(1) * 1000.0f);
    QSDEW = 0.0f;
    QSEVA = 0.0f;
    if (soilice(1) < 0.0f) {
      c.soilwat(1) = c.soilwat(1) + soilice(1);
      soilice(1) = 0.0f;
    }
  }

  qinsrf = (PONDING + c.PONDING1 + c.PONDING2) / DT * 0.001f;   // func:4757
  if (c.ISNOW == 0) qinsrf = qinsrf + (c.QSNBOT + QSDEW + QRAIN) * 0.001f;  // func:4760-4764
  else qinsrf = qinsrf + (c.QSNBOT + QSDEW) * 0.001f;

  QSEVA = QSEVA * 0.001f;                                       // func:4766

  const int NROOT = t.LK_NROOT[c.lutyp];
  for (int IZ = 1; IZ <= NROOT; ++IZ) ETRANI(IZ) = c.ETRAN * BTRANI(IZ) * 0.001f;  // func:4768-4770

  if (c.IST == 2) {                                             // func:4774-4777
    c.runsrf = 0.0f;
    if (c.WSLAKE >= WSLMAX) c.runsrf = qinsrf * 1000.0f;
    c.WSLAKE = c.WSLAKE + (qinsrf - QSEVA) * 1000.0f * DT - c.runsrf * DT;
  } else {
    float QDRAIN = 0.0f, FCRMAX = 0.0f;
    soilh2o(cx, c.sltyp, c.lutyp, DT, c.ZSOIL, DZSNSO, c.slptyp, qinsrf, QSEVA, ETRANI, soilice,
            c.soilwat, c.SMC, c.ZWT, c.runsrf, QDRAIN, c.runsub, WCND, FCRMAX, c.niter_soil);  // func:4779-4783

    if (cx.opt.opt_run == 1) {                                  // func:4785-4791
      groundwater(cx, c.sltyp, DT, c.ZSOIL, soilice, WCND, FCRMAX, c.soilwat, c.ZWT, c.WA, c.WT, QIN,
                  QDIS);
      c.runsub = QDIS;
    }
    if (cx.opt.opt_run == 3 || cx.opt.opt_run == 4) c.runsub = c.runsub + QDRAIN;  // func:4793-4795

    for (int IZ = 1; IZ <= NSOIL; ++IZ) c.SMC(IZ) = c.soilwat(IZ) + soilice(IZ);   // func:4797-4799
  }

  c.runsub = c.runsub + SNOFLOW;                                // func:4802
}

// func:4807-5046
void canwater(const Ctx& cx, int lutyp, float DT, float SFCTMP, float UU, float VV, float FCEV,
              float FCTR, float QPRECC, float QPRECL, float elai, float esai, int IST, float TG,
              float fveg, bool FROZEN_CANOPY, float& CANLIQ, float& CANICE, float& TV, float& CMC,
              float& ECAN, float& ETRAN, float& QRAIN, float& QSNOW, float& snowhin, float& FWET,
              float& FPICE) {
  const Tables& t = cx.tbl;
  float FP = 0.0f, RAIN = 0.0f, SNOW = 0.0f;                    // func:4875-4888
  float QINTR = 0.0f, QDRIPR = 0.0f, QTHROR = 0.0f, QINTS = 0.0f, QDRIPS = 0.0f, QTHROS = 0.0f;
  QRAIN = 0.0f;
  QSNOW = 0.0f;
  snowhin = 0.0f;
  ECAN = 0.0f;
  FPICE = 0.0f;

  if (cx.opt.opt_snf == 1) {                                    // func:4893-4905
    if (SFCTMP > TFRZ + 2.5f) {
      FPICE = 0.0f;
    } else {
      if (SFCTMP <= TFRZ + 0.5f) FPICE = 1.0f;
      else if (SFCTMP <= TFRZ + 2.0f) FPICE = 1.0f - (-54.632f + 0.2f * SFCTMP);
      else FPICE = 0.6f;
    }
  }
  if (cx.opt.opt_snf == 2) {                                    // func:4907-4913
    if (SFCTMP >= TFRZ + 2.2f) FPICE = 0.0f;
    else FPICE = 1.0f;
  }
  if (cx.opt.opt_snf == 3) {                                    // func:4915-4921
    if (SFCTMP >= TFRZ) FPICE = 0.0f;
    else FPICE = 1.0f;
  }

  float BDFALL = fmin2(120.0f, 67.92f + 51.25f * expf((SFCTMP - TFRZ) / 2.59f));  // func:4926

  RAIN = (QPRECC + QPRECL) * (1.0f - FPICE);                    // func:4928-4929
  SNOW = (QPRECC + QPRECL) * FPICE;

  if (QPRECC + QPRECL > 0.0f) FP = (QPRECC + QPRECL) / (10.0f * QPRECC + QPRECL);  // func:4933-4934

  float MAXLIQ = t.LK_CANWMXP[lutyp] * (elai + esai);           // func:4939

  if ((elai + esai) > 0.0f) {                                   // func:4943-4953
    QINTR = fveg * RAIN * FP;
    QINTR = fmin2(QINTR, (MAXLIQ - CANLIQ) / DT * (1.0f - expf(-RAIN * DT / MAXLIQ)));
    QINTR = fmax2(QINTR, 0.0f);
    QDRIPR = fveg * RAIN - QINTR;
    QTHROR = (1.0f - fveg) * RAIN;
  } else {
    QINTR = 0.0f;
    QDRIPR = 0.0f;
    QTHROR = RAIN;
  }

  float QEVAC, QDEWC, QSUBC, QFROC;
  if (!FROZEN_CANOPY) {                                         // func:4956-4968
    ETRAN = fmax2(FCTR / HVAP, 0.0f);
    QEVAC = fmax2(FCEV / HVAP, 0.0f);
    QDEWC = fabsf(fmin2(FCEV / HVAP, 0.0f));
    QSUBC = 0.0f;
    QFROC = 0.0f;
  } else {
    ETRAN = fmax2(FCTR / HSUB, 0.0f);
    QEVAC = 0.0f;
    QDEWC = 0.0f;
    QSUBC = fmax2(FCEV / HSUB, 0.0f);
    QFROC = fabsf(fmin2(FCEV / HSUB, 0.0f));
  }

  QEVAC = fmin2(CANLIQ / DT, QEVAC);                            // func:4973-4975
  CANLIQ = fmax2(0.0f, CANLIQ + (QINTR + QDEWC - QEVAC) * DT);
  if (CANLIQ <= 1.0E-6f) CANLIQ = 0.0f;

  float MAXSNO = 6.6f * (0.27f + 46.0f / BDFALL) * (elai + esai);  // func:4978

  if ((elai + esai) > 0.0f) {                                   // func:4980-4992
    QINTS = fveg * SNOW * FP;
    QINTS = fmin2(QINTS, (MAXSNO - CANICE) / DT * (1.0f - expf(-SNOW * DT / MAXSNO)));
    QINTS = fmax2(QINTS, 0.0f);
    float FT = fmax2(0.0f, (TV - 270.15f) / 1.87E5f);
    float FV = sqrtf(UU * UU + VV * VV) / 1.56E5f;
    QDRIPS = fmax2(0.0f, CANICE) * (FV + FT);
    QTHROS = (1.0f - fveg) * SNOW + (fveg * SNOW - QINTS);
  } else {
    QINTS = 0.0f;
    QDRIPS = 0.0f;
    QTHROS = SNOW;
  }

  QSUBC = fmin2(CANICE / DT, QSUBC);                            // func:4994-4996
  CANICE = fmax2(0.0f, CANICE + (QINTS - QDRIPS) * DT + (QFROC - QSUBC) * DT);
  if (CANICE <= 1.0E-6f) CANICE = 0.0f;

  if (CANICE > 0.0f) FWET = fmax2(0.0f, CANICE) / fmax2(MAXSNO, 1.0E-06f);  // func:5000-5005
  else FWET = fmax2(0.0f, CANLIQ) / fmax2(MAXLIQ, 1.0E-06f);
  FWET = powf(fmin2(FWET, 1.0f), 0.667f);

  float QMELTC = 0.0f, QFRZC = 0.0f;                            // func:5009-5010

  if (CANICE > 1.0E-6f && TV > TFRZ) {                          // func:5012-5017
    QMELTC = fmin2(CANICE / DT, (TV - TFRZ) * CICE * CANICE / DENICE / (DT * HFUS));
    CANICE = fmax2(0.0f, CANICE - QMELTC * DT);
    CANLIQ = fmax2(0.0f, CANLIQ + QMELTC * DT);
    TV = FWET * TFRZ + (1.0f - FWET) * TV;
  }

  if (CANLIQ > 1.0E-6f && TV < TFRZ) {                          // func:5019-5024
    QFRZC = fmin2(CANLIQ / DT, (TFRZ - TV) * CWAT * CANLIQ / DENWAT / (DT * HFUS));
    CANLIQ = fmax2(0.0f, CANLIQ - QFRZC * DT);
    CANICE = fmax2(0.0f, CANICE + QFRZC * DT);
    TV = FWET * TFRZ + (1.0f - FWET) * TV;
  }

  CMC = CANLIQ + CANICE;                                        // func:5028
  ECAN = QEVAC + QSUBC - QDEWC - QFROC;                         // func:5032
  QRAIN = QDRIPR + QTHROR;                                      // func:5036-5038
  QSNOW = QDRIPS + QTHROS;
  snowhin = QSNOW / BDFALL;

  if (IST == 2 && TG > TFRZ) {                                  // func:5041-5044
    QSNOW = 0.0f;
    snowhin = 0.0f;
  }
}

// func:5049-5174
void snowwater(const Ctx& cx, float DT, const A4& ZSOIL, const I7& IMELT, float SFCTMP, float snowhin,
               float QSNOW, float QSNFRO, float QSNSUB, float QRAIN, const A3& FICEOLD, int& ISNOW,
               float& snowh, float& sneqv, A3& SNICE, A3& SNLIQ, A4& soilwat, A4& soilice, A7& STC,
               A7& ZSNSO, A7& DZSNSO, float& QSNBOT, float& SNOFLOW, float& PONDING1,
               float& PONDING2) {
  SNOFLOW = 0.0f;                                               // func:5093-5095
  PONDING1 = 0.0f;
  PONDING2 = 0.0f;

  snowfall(DT, QSNOW, snowhin, SFCTMP, ISNOW, snowh, DZSNSO, STC, SNICE, SNLIQ, sneqv);  // func:5097-5100

  if (ISNOW < 0) compact(DT, STC, SNICE, SNLIQ, IMELT, FICEOLD, ISNOW, DZSNSO);          // func:5104-5107
  if (ISNOW < 0) combine(ISNOW, soilwat, STC, SNICE, SNLIQ, DZSNSO, soilice, snowh, sneqv, PONDING1,
                         PONDING2);                                                      // func:5109-5113
  if (ISNOW < 0) divide(ISNOW, STC, SNICE, SNLIQ, DZSNSO);                               // func:5115-5117

  snowh2o(cx, DT, QSNFRO, QSNSUB, QRAIN, ISNOW, DZSNSO, snowh, sneqv, SNICE, SNLIQ, soilwat, soilice,
          STC, QSNBOT, PONDING1, PONDING2);                     // func:5119-5123

  for (int iz = -NSNOW + 1; iz <= ISNOW; ++iz) {                // func:5127-5133
    SNICE(iz) = 0.0f;
    SNLIQ(iz) = 0.0f;
    STC(iz) = 0.0f;
    DZSNSO(iz) = 0.0f;
    ZSNSO(iz) = 0.0f;
  }

  if (sneqv > 2000.0f) {                                        // func:5137-5143
    float BDSNOW = SNICE(0) / DZSNSO(0);
    SNOFLOW = (sneqv - 2000.0f);
    SNICE(0) = SNICE(0) - SNOFLOW;
    DZSNSO(0) = DZSNSO(0) - SNOFLOW / BDSNOW;
    SNOFLOW = SNOFLOW / DT;
  }

  if (ISNOW < 0) {                                              // func:5147-5152
    sneqv = 0.0f;
    for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) sneqv = sneqv + SNICE(IZ) + SNLIQ(IZ);
  }

  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) DZSNSO(IZ) = -DZSNSO(IZ);  // func:5156-5158

  DZSNSO(1) = ZSOIL(1);                                         // func:5160-5163
  for (int IZ = 2; IZ <= NSOIL; ++IZ) DZSNSO(IZ) = (ZSOIL(IZ) - ZSOIL(IZ - 1));

  ZSNSO(ISNOW + 1) = DZSNSO(ISNOW + 1);                         // func:5165-5168
  for (int IZ = ISNOW + 2; IZ <= NSOIL; ++IZ) ZSNSO(IZ) = ZSNSO(IZ - 1) + DZSNSO(IZ);

  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) DZSNSO(IZ) = -DZSNSO(IZ);  // func:5170-5172
}

// func:5177-5233
void snowfall(float DT, float QSNOW, float snowhin, float SFCTMP, int& ISNOW, float& snowh,
              A7& DZSNSO, A7& STC, A3& SNICE, A3& SNLIQ, float& sneqv) {
  int NEWNODE = 0;
  if (ISNOW == 0 && QSNOW > 0.0f) {                             // func:5209-5212
    snowh = snowh + snowhin * DT;
    sneqv = sneqv + QSNOW * DT;
  }
  if (ISNOW == 0 && QSNOW > 0.0f && snowh >= 0.025f) {          // func:5216-5225
    ISNOW = -1;
    NEWNODE = 1;
    DZSNSO(0) = snowh;
    snowh = 0.0f;
    STC(0) = fmin2(273.16f, SFCTMP);
    SNICE(0) = sneqv;
    SNLIQ(0) = 0.0f;
  }
  if (ISNOW < 0 && NEWNODE == 0 && QSNOW > 0.0f) {              // func:5229-5232
    SNICE(ISNOW + 1) = SNICE(ISNOW + 1) + QSNOW * DT;
    DZSNSO(ISNOW + 1) = DZSNSO(ISNOW + 1) + snowhin * DT;
  }
}

// func:5536-5577
void combo(float& DZ, float& WLIQ, float& WICE, float& T, float DZ2, float WLIQ2, float WICE2, float T2) {
  float DZC = DZ + DZ2;
  float WICEC = WICE + WICE2;
  float WLIQC = WLIQ + WLIQ2;
  float H = (CICE * WICE + CWAT * WLIQ) * (T - TFRZ) + HFUS * WLIQ;
  float H2 = (CICE * WICE2 + CWAT * WLIQ2) * (T2 - TFRZ) + HFUS * WLIQ2;
  float HC = H + H2;
  float TC;
  if (HC < 0.0f) TC = TFRZ + HC / (CICE * WICEC + CWAT * WLIQC);
  else if (HC <= HFUS * WLIQC) TC = TFRZ;
  else TC = TFRZ + (HC - HFUS * WLIQC) / (CICE * WICEC + CWAT * WLIQC);
  DZ = DZC;
  WICE = WICEC;
  WLIQ = WLIQC;
  T = TC;
}

// func:5236-5413
void combine(int& ISNOW, A4& soilwat, A7& STC, A3& SNICE, A3& SNLIQ, A7& DZSNSO, A4& soilice,
             float& snowh, float& sneqv, float& PONDING1, float& PONDING2) {
  const float DZMIN[4] = {0.0f, 0.025f, 0.025f, 0.1f};          // func:5272 (1-based)
  int ISNOW_OLD = ISNOW;                                        // func:5274

  for (int J = ISNOW_OLD + 1; J <= 0; ++J) {                    // func:5276-5318
    if (SNICE(J) <= 0.1f) {
      if (J != 0) {
        SNLIQ(J + 1) = SNLIQ(J + 1) + SNLIQ(J);
        SNICE(J + 1) = SNICE(J + 1) + SNICE(J);
      } else {
        if (ISNOW_OLD < -1) {
          SNLIQ(J - 1) = SNLIQ(J - 1) + SNLIQ(J);
          SNICE(J - 1) = SNICE(J - 1) + SNICE(J);
        } else {
          if (SNICE(J) >= 0.0f) {
            PONDING1 = SNLIQ(J);
            sneqv = SNICE(J);
            snowh = DZSNSO(J);
          } else {
            PONDING1 = SNLIQ(J) + SNICE(J);
            if (PONDING1 < 0.0f) {
              soilice(1) = fmax2(0.0f, soilice(1) + PONDING1 / (DZSNSO(1) * 1000.0f));
              PONDING1 = 0.0f;
            }
            sneqv = 0.0f;
            snowh = 0.0f;
          }
          SNLIQ(J) = 0.0f;
          SNICE(J) = 0.0f;
          DZSNSO(J) = 0.0f;
        }
      }
      if (J > ISNOW + 1 && ISNOW < -1) {                        // func:5308-5315
        for (int I = J; I >= ISNOW + 2; --I) {
          STC(I) = STC(I - 1);
          SNLIQ(I) = SNLIQ(I - 1);
          SNICE(I) = SNICE(I - 1);
          DZSNSO(I) = DZSNSO(I - 1);
        }
      }
      ISNOW = ISNOW + 1;                                        // func:5316
    }
  }

  if (soilice(1) < 0.0f) {                                      // func:5322-5325
    soilwat(1) = soilwat(1) + soilice(1);
    soilice(1) = 0.0f;
  }

  if (ISNOW == 0) return;                                       // func:5327

  sneqv = 0.0f;                                                 // func:5329-5339
  snowh = 0.0f;
  float ZWICE = 0.0f;
  float ZWLIQ = 0.0f;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    sneqv = sneqv + SNICE(J) + SNLIQ(J);
    snowh = snowh + DZSNSO(J);
    ZWICE = ZWICE + SNICE(J);
    ZWLIQ = ZWLIQ + SNLIQ(J);
  }

  if (snowh < 0.025f && ISNOW < 0) {                            // func:5344-5350
    ISNOW = 0;
    sneqv = ZWICE;
    PONDING2 = ZWLIQ;
    if (sneqv <= 0.0f) snowh = 0.0f;
  }

  if (ISNOW < -1) {                                             // func:5361-5411
    ISNOW_OLD = ISNOW;
    int MSSI = 1;
    for (int I = ISNOW_OLD + 1; I <= 0; ++I) {
      if (DZSNSO(I) < DZMIN[MSSI]) {
        int NEIBOR;
        if (I == ISNOW + 1) {
          NEIBOR = I + 1;
        } else if (I == 0) {
          NEIBOR = I - 1;
        } else {
          NEIBOR = I + 1;
          if ((DZSNSO(I - 1) + DZSNSO(I)) < (DZSNSO(I + 1) + DZSNSO(I))) NEIBOR = I - 1;
        }
        int J, L;
        if (NEIBOR > I) {
          J = NEIBOR;
          L = I;
        } else {
          J = I;
          L = NEIBOR;
        }
        combo(DZSNSO(J), SNLIQ(J), SNICE(J), STC(J), DZSNSO(L), SNLIQ(L), SNICE(L), STC(L));  // func:5387-5388

        if (J - 1 > ISNOW + 1) {                                // func:5391-5398
          for (int K = J - 1; K >= ISNOW + 2; --K) {
            STC(K) = STC(K - 1);
            SNICE(K) = SNICE(K - 1);
            SNLIQ(K) = SNLIQ(K - 1);
            DZSNSO(K) = DZSNSO(K - 1);
          }
        }
        ISNOW = ISNOW + 1;                                      // func:5401-5402
        if (ISNOW >= -1) break;
      } else {
        MSSI = MSSI + 1;                                        // func:5406
      }
    }
  }
}

// func:5416-5533
void divide(int& ISNOW, A7& STC, A3& SNICE, A3& SNLIQ, A7& DZSNSO) {
  float DZ[4] = {0, 0, 0, 0}, SWICE[4] = {0, 0, 0, 0}, SWLIQ[4] = {0, 0, 0, 0}, TSNO[4] = {0, 0, 0, 0};
  const int absI = ISNOW < 0 ? -ISNOW : ISNOW;
  for (int J = 1; J <= NSNOW; ++J) {                            // func:5443-5450
    if (J <= absI) {
      DZ[J] = DZSNSO(J + ISNOW);
      SWICE[J] = SNICE(J + ISNOW);
      SWLIQ[J] = SNLIQ(J + ISNOW);
      TSNO[J] = STC(J + ISNOW);
    }
  }
  int MSNO = absI;                                              // func:5452

  if (MSNO == 1) {                                              // func:5454-5466
    if (DZ[1] > 0.05f) {
      MSNO = 2;
      DZ[1] = DZ[1] / 2.0f;
      SWICE[1] = SWICE[1] / 2.0f;
      SWLIQ[1] = SWLIQ[1] / 2.0f;
      DZ[2] = DZ[1];
      SWICE[2] = SWICE[1];
      SWLIQ[2] = SWLIQ[1];
      TSNO[2] = TSNO[1];
    }
  }

  if (MSNO > 1) {                                               // func:5468-5502
    if (DZ[1] > 0.05f) {
      float DRR = DZ[1] - 0.05f;
      float PROPOR = DRR / DZ[1];
      float ZWICE = PROPOR * SWICE[1];
      float ZWLIQ = PROPOR * SWLIQ[1];
      PROPOR = 0.05f / DZ[1];
      SWICE[1] = PROPOR * SWICE[1];
      SWLIQ[1] = PROPOR * SWLIQ[1];
      DZ[1] = 0.05f;

      combo(DZ[2], SWLIQ[2], SWICE[2], TSNO[2], DRR, ZWLIQ, ZWICE, TSNO[1]);  // func:5479-5480

      if (MSNO <= 2 && DZ[2] > 0.20f) {                         // func:5483-5499
        MSNO = 3;
        float DTDZ = (TSNO[1] - TSNO[2]) / ((DZ[1] + DZ[2]) / 2.0f);
        DZ[2] = DZ[2] / 2.0f;
        SWICE[2] = SWICE[2] / 2.0f;
        SWLIQ[2] = SWLIQ[2] / 2.0f;
        DZ[3] = DZ[2];
        SWICE[3] = SWICE[2];
        SWLIQ[3] = SWLIQ[2];
        TSNO[3] = TSNO[2] - DTDZ * DZ[2] / 2.0f;
        if (TSNO[3] >= TFRZ) {
          TSNO[3] = TSNO[2];
        } else {
          TSNO[2] = TSNO[2] + DTDZ * DZ[2] / 2.0f;
        }
      }
    }
  }

  if (MSNO > 2) {                                               // func:5504-5517
    if (DZ[2] > 0.2f) {
      float DRR = DZ[2] - 0.2f;
      float PROPOR = DRR / DZ[2];
      float ZWICE = PROPOR * SWICE[2];
      float ZWLIQ = PROPOR * SWLIQ[2];
      PROPOR = 0.2f / DZ[2];
      SWICE[2] = PROPOR * SWICE[2];
      SWLIQ[2] = PROPOR * SWLIQ[2];
      DZ[2] = 0.2f;
      combo(DZ[3], SWLIQ[3], SWICE[3], TSNO[3], DRR, ZWLIQ, ZWICE, TSNO[2]);
    }
  }

  ISNOW = -MSNO;                                                // func:5519

  for (int J = ISNOW + 1; J <= 0; ++J) {                        // func:5521-5526
    DZSNSO(J) = DZ[J - ISNOW];
    SNICE(J) = SWICE[J - ISNOW];
    SNLIQ(J) = SWLIQ[J - ISNOW];
    STC(J) = TSNO[J - ISNOW];
  }
}

// func:5580-5677
void compact(float DT, const A7& STC, const A3& SNICE, const A3& SNLIQ, const I7& IMELT,
             const A3& FICEOLD, int ISNOW, A7& DZSNSO) {
  const float C2 = 21.e-3f, C3 = 2.5e-6f, C4 = 0.04f, C5 = 2.0f, DM = 100.0f, ETA0 = 0.8e+6f;
  A3 FICE;
  float BURDEN = 0.0f;                                          // func:5624
  for (int J = ISNOW + 1; J <= 0; ++J) {                        // func:5626-5675
    float WX = SNICE(J) + SNLIQ(J);
    FICE(J) = SNICE(J) / WX;
    float VOID = 1.0f - (SNICE(J) / DENICE + SNLIQ(J) / DENWAT) / DZSNSO(J);

    if (VOID > 0.001f && SNICE(J) > 0.1f) {                     // func:5633
      float BI = SNICE(J) / DZSNSO(J);
      float TD = fmax2(0.0f, TFRZ - STC(J));
      float DEXPF = expf(-C4 * TD);

      float DDZ1 = -C3 * DEXPF;                                 // func:5640
      if (BI > DM) DDZ1 = DDZ1 * expf(-46.0E-3f * (BI - DM));   // func:5642
      if (SNLIQ(J) > 0.01f * DZSNSO(J)) DDZ1 = DDZ1 * C5;       // func:5646

      float DDZ2 = -(BURDEN + 0.5f * WX) * expf(-0.08f * TD - C2 * BI) / ETA0;  // func:5650

      float DDZ3;
      if (IMELT(J) == 1) {                                      // func:5654-5659
        DDZ3 = fmax2(0.0f, (FICEOLD(J) - FICE(J)) / fmax2(1.E-6f, FICEOLD(J)));
        DDZ3 = -DDZ3 / DT;
      } else {
        DDZ3 = 0.0f;
      }

      float PDZDTC = (DDZ1 + DDZ2 + DDZ3) * DT;                 // func:5663-5664
      PDZDTC = fmax2(-0.5f, PDZDTC);

      DZSNSO(J) = DZSNSO(J) * (1.0f + PDZDTC);                  // func:5668
    }
    BURDEN = BURDEN + WX;                                       // func:5673
  }
}

// func:5680-5819
void snowh2o(const Ctx& cx, float DT, float QSNFRO, float QSNSUB, float QRAIN, int& ISNOW,
             A7& DZSNSO, float& snowh, float& sneqv, A3& SNICE, A3& SNLIQ, A4& soilwat, A4& soilice,
             A7& STC, float& QSNBOT, float& PONDING1, float& PONDING2) {
  A3 VOL_LIQ, VOL_ICE, epore;
  for (int i = 0; i < 3; ++i) { VOL_LIQ.v[i] = 0.0f; VOL_ICE.v[i] = 0.0f; epore.v[i] = 0.0f; }

  if (sneqv == 0.0f) {                                          // func:5726-5732
    soilice(1) = soilice(1) + (QSNFRO - QSNSUB) * DT / (DZSNSO(1) * 1000.0f);
    if (soilice(1) < 0.0f) {
      soilwat(1) = soilwat(1) + soilice(1);
      soilice(1) = 0.0f;
    }
  }

  if (ISNOW == 0 && sneqv > 0.0f) {                             // func:5739-5754
    float TEMP = sneqv;
    sneqv = sneqv - QSNSUB * DT + QSNFRO * DT;
    float PROPOR = sneqv / TEMP;
    snowh = fmax2(0.0f, PROPOR * snowh);

    if (sneqv < 0.0f) {
      soilice(1) = soilice(1) + sneqv / (DZSNSO(1) * 1000.0f);
      sneqv = 0.0f;
      snowh = 0.0f;
    }
    if (soilice(1) < 0.0f) {
      soilwat(1) = soilwat(1) + soilice(1);
      soilice(1) = 0.0f;
    }
  }

  if (snowh <= 1.0E-8f || sneqv <= 1.0E-6f) {                   // func:5756-5759
    snowh = 0.0f;
    sneqv = 0.0f;
  }

  if (ISNOW < 0) {                                              // func:5763-5778
    float WGDIF = SNICE(ISNOW + 1) - QSNSUB * DT + QSNFRO * DT;
    SNICE(ISNOW + 1) = WGDIF;
    if (WGDIF < 1.0E-6f && ISNOW < 0) {
      combine(ISNOW, soilwat, STC, SNICE, SNLIQ, DZSNSO, soilice, snowh, sneqv, PONDING1, PONDING2);
    }
    if (ISNOW < 0) {
      SNLIQ(ISNOW + 1) = SNLIQ(ISNOW + 1) + QRAIN * DT;
      SNLIQ(ISNOW + 1) = fmax2(0.0f, SNLIQ(ISNOW + 1));
    }
  }

  for (int J = -NSNOW + 1; J <= 0; ++J) {                       // func:5784-5790
    if (J >= ISNOW + 1) {
      VOL_ICE(J) = fmin2(1.0f, SNICE(J) / (DZSNSO(J) * DENICE));
      epore(J) = 1.0f - VOL_ICE(J);
      VOL_LIQ(J) = fmin2(epore(J), SNLIQ(J) / (DZSNSO(J) * DENWAT));
    }
  }

  float QIN = 0.0f;                                             // func:5792-5793
  float QOUT = 0.0f;

  for (int J = -NSNOW + 1; J <= 0; ++J) {                       // func:5797-5814
    if (J >= ISNOW + 1) {
      SNLIQ(J) = SNLIQ(J) + QIN;
      if (J <= -1) {
        if (epore(J) < 0.05f || epore(J + 1) < 0.05f) {
          QOUT = 0.0f;
        } else {
          QOUT = fmax2(0.0f, (VOL_LIQ(J) - cx.tbl.KK_SSI * epore(J)) * DZSNSO(J));
          QOUT = fmin2(QOUT, (1.0f - VOL_ICE(J + 1) - VOL_LIQ(J + 1)) * DZSNSO(J + 1));
        }
      } else {
        QOUT = fmax2(0.0f, (VOL_LIQ(J) - cx.tbl.KK_SSI * epore(J)) * DZSNSO(J));
      }
      QOUT = QOUT * 1000.0f;
      SNLIQ(J) = SNLIQ(J) - QOUT;
      QIN = QOUT;
    }
  }

  QSNBOT = QOUT / DT;                                           // func:5818
}

}  // namespace noahmp_oracle
