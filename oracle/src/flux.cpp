// TEST INFRASTRUCTURE ONLY (see common.h).
// vege_flux, bare_flux, ragrb, sfcdif1, sfcdif2, esat, stomata, canres, calhum :
// core/module_noahmp_func.f90:2465-3984
#include "physics.h"

namespace noahmp_oracle {

static inline float TDC(float T) { return fmin2(50.0f, fmax2(-50.0f, (T - TFRZ))); }  // func:2688,3120

// func:3692-3736
void esat(float T, float& ESW, float& ESI, float& DESW, float& DESI) {
  const float A0 = 6.107799961f, A1 = 4.436518521E-01f, A2_ = 1.428945805E-02f, A3_ = 2.650648471E-04f,
              A4_ = 3.031240396E-06f, A5 = 2.034080948E-08f, A6 = 6.136820929E-11f;
  const float B0 = 6.109177956f, B1 = 5.034698970E-01f, B2 = 1.886013408E-02f, B3 = 4.176223716E-04f,
              B4 = 5.824720280E-06f, B5 = 4.838803174E-08f, B6 = 1.838826904E-10f;
  const float C0 = 4.438099984E-01f, C1 = 2.857002636E-02f, C2 = 7.938054040E-04f, C3 = 1.215215065E-05f,
              C4 = 1.036561403E-07f, C5 = 3.532421810e-10f, C6 = -7.090244804E-13f;
  const float D0 = 5.030305237E-01f, D1 = 3.773255020E-02f, D2 = 1.267995369E-03f, D3 = 2.477563108E-05f,
              D4 = 3.005693132E-07f, D5 = 2.158542548E-09f, D6 = 7.131097725E-12f;
  ESW = 100.0f * (A0 + T * (A1 + T * (A2_ + T * (A3_ + T * (A4_ + T * (A5 + T * A6))))));
  ESI = 100.0f * (B0 + T * (B1 + T * (B2 + T * (B3 + T * (B4 + T * (B5 + T * B6))))));
  DESW = 100.0f * (C0 + T * (C1 + T * (C2 + T * (C3 + T * (C4 + T * (C5 + T * C6))))));
  DESI = 100.0f * (D0 + T * (D1 + T * (D2 + T * (D3 + T * (D4 + T * (D5 + T * D6))))));
}

// func:2465-2964
void vege_flux(const Ctx& cx, ColumnIO& c, int ISNOW, int lutyp, float DT, float SAV, float SAG,
               float LWDN, float UR, float UU, float VV, float SFCTMP, float THAIR, float QAIR,
               float EAIR, float RHOAIR, float snowh, float VAI, float GAMMAV, float GAMMAG,
               float FWET, float LAISUN, float LAISHA, float CWP, const A7& DZSNSO, float HTOP,
               float ZLVL, float ZPD, float Z0M, float fveg, float Z0MG, float EMV, float EMG,
               float CANLIQ, float CANICE, const A7& STC, const A7& DF, float& RSSUN, float& RSSHA,
               float RSURF, float LATHEAV, float LATHEAG, float PARSUN, float PARSHA, float IGS,
               float FOLN, float CO2AIR, float O2AIR, float BTRAN, float SFCPRS, float RHSUR,
               float& EAH, float& TAH, float& TV, float& TG, float& CM, float& CH, float& TAUXV,
               float& TAUYV, float& IRG, float& IRC, float& SHG, float& SHC, float& EVG, float& EVC,
               float& TR, float& GH, float& T2MV, float& PSNSUN, float& PSNSHA, float& QSFC,
               float PSFC, float& Q2V, float& CAH2, float& CHLEAF, float& CHUC) {
  (void)LATHEAG;
  const int NITERC = 20;                                        // func:2675
  const int NITERG = 5;                                         // func:2677
  float MPE_ = 1E-6f;                                           // func:2690
  int LITER = 0;
  float FV = 0.1f;
  float DTV = 0.0f, DTG = 0.0f;                                 // func:2695-2701
  int MOZSGN = 0;
  float HG = 0.0f, H = 0.0f;
  // locals the Fortran leaves undefined before first use are zero-initialised here
  float MOZ = 0.0f, FM = 0.0f, FH = 0.0f, FM2 = 0.0f, FH2 = 0.0f, CH2 = 0.0f, MOZG = 0.0f, FHG = 0.0f;
  float WSTAR = 0.0f;
  float RAMG = 0.0f, RAHG = 0.0f, RAWG = 0.0f, RB = 0.0f;
  float Z0H = 0.0f, Z0HG = 0.0f;
  float ESATW, ESATI, DSATW, DSATI, ESTG, ESTV = 0.0f, DESTV = 0.0f, DESTG = 0.0f;
  float CAH = 0.0f, CVH = 0.0f, CGH, CSH, CEV, CIR, AIR;

  float VAIE = fmin2(6.0f, VAI / fveg);                         // func:2704-2706
  float LAISUNE = fmin2(6.0f, LAISUN / fveg);
  float LAISHAE = fmin2(6.0f, LAISHA / fveg);

  float T = TDC(TG);                                            // func:2710-2716
  esat(T, ESATW, ESATI, DSATW, DSATI);
  if (T > 0.0f) ESTG = ESATW; else ESTG = ESATI;

  QSFC = 0.622f * EAIR / (PSFC - 0.378f * EAIR);                // func:2720

  float HCAN = HTOP;                                            // func:2724-2725
  float UC = UR * logf(HCAN / Z0M) / logf(ZLVL / Z0M);
  if ((HCAN - ZPD) <= 0.0f) c.flags |= FLAG_HCAN;               // func:2726-2738

  AIR = -EMV * (1.0f + (1.0f - EMV) * (1.0f - EMG)) * LWDN - EMV * EMG * SB * pow4i(TG);  // func:2741
  CIR = (2.0f - EMV * (1.0f - EMG)) * EMV * SB;                                            // func:2742

  int iters_done = 0;
  for (int iter = 1; iter <= NITERC; ++iter) {                  // func:2744 loop1
    iters_done = iter;
    Z0H = Z0M;                                                  // func:2746-2752
    Z0HG = Z0MG;

    if (cx.opt.opt_sfc == 1) {                                  // func:2755-2761
      sfcdif1(iter, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, Z0H, UR, MPE_, MOZ, MOZSGN, FM, FH, FM2,
              FH2, CM, CH, FV, CH2, c.flags);
    }
    if (cx.opt.opt_sfc == 2) {                                  // func:2763-2772
      sfcdif2(iter, Z0M, TAH, THAIR, UR, cx.tbl.KK_CZIL, ZLVL, CM, CH, MOZ, WSTAR, FV);
      CH = CH / UR;
      CM = CM / UR;
    }

    float RAMC = fmax2(1.0f, 1.0f / (CM * UR));                 // func:2774-2776
    float RAHC = fmax2(1.0f, 1.0f / (CH * UR));
    float RAWC = RAHC;
    (void)RAMC;

    ragrb(cx, iter, VAIE, RHOAIR, HG, TAH, ZPD, Z0MG, Z0HG, HCAN, UC, Z0H, FV, CWP, lutyp, MPE_, TV,
          MOZG, FHG, RAMG, RAHG, RAWG, RB);                     // func:2780-2784

    T = TDC(TV);                                                // func:2787-2795
    esat(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.0f) { ESTV = ESATW; DESTV = DSATW; }
    else { ESTV = ESATI; DESTV = DSATI; }

    if (iter == 1) {                                            // func:2798-2814
      if (cx.opt.opt_crs == 1) {
        stomata(cx, lutyp, IGS, SFCPRS, SFCTMP, PARSUN, TV, EAH, ESTV, O2AIR, CO2AIR, FOLN, BTRAN, RB,
                RSSUN, PSNSUN, c.iter_stom_sun);
        stomata(cx, lutyp, IGS, SFCPRS, SFCTMP, PARSHA, TV, EAH, ESTV, O2AIR, CO2AIR, FOLN, BTRAN, RB,
                RSSHA, PSNSHA, c.iter_stom_sha);
      }
      if (cx.opt.opt_crs == 2) {
        canres(cx, lutyp, SFCPRS, TV, PARSUN, EAH, BTRAN, RSSUN, PSNSUN);
        canres(cx, lutyp, SFCPRS, TV, PARSHA, EAH, BTRAN, RSSHA, PSNSHA);
      }
    }

    CAH = 1.0f / RAHC;                                          // func:2817-2823
    CVH = 2.0f * VAIE / RB;
    CGH = 1.0f / RAHG;
    float COND = CAH + CVH + CGH;
    float ATA = (SFCTMP * CAH + TG * CGH) / COND;
    float BTA = CVH / COND;
    CSH = (1.0f - BTA) * RHOAIR * CPAIR * CVH;

    float CAW = 1.0f / RAWC;                                    // func:2826-2834
    float CEW = FWET * VAIE / RB;
    float CTW = (1.0f - FWET) * (LAISUNE / (RB + RSSUN) + LAISHAE / (RB + RSSHA));
    float CGW = 1.0f / (RAWG + RSURF);
    COND = CAW + CEW + CTW + CGW;
    float AEA = (EAIR * CAW + ESTG * CGW) / COND;
    float BEA = (CEW + CTW) / COND;
    CEV = (1.0f - BEA) * CEW * RHOAIR * CPAIR / GAMMAV;
    float CTR = (1.0f - BEA) * CTW * RHOAIR * CPAIR / GAMMAV;

    TAH = ATA + BTA * TV;                                       // func:2837-2838
    EAH = AEA + BEA * ESTV;

    IRC = fveg * (AIR + CIR * pow4i(TV));                       // func:2840-2848
    SHC = fveg * RHOAIR * CPAIR * CVH * (TV - TAH);
    EVC = fveg * RHOAIR * CPAIR * CEW * (ESTV - EAH) / GAMMAV;
    TR = fveg * RHOAIR * CPAIR * CTW * (ESTV - EAH) / GAMMAV;
    if (TV > TFRZ) EVC = fmin2(CANLIQ * LATHEAV / DT, EVC);
    else EVC = fmin2(CANICE * LATHEAV / DT, EVC);

    float B = SAV - IRC - SHC - EVC - TR;                       // func:2850-2852
    float A = fveg * (4.0f * CIR * pow3i(TV) + CSH + (CEV + CTR) * DESTV);
    DTV = B / A;

    IRC = IRC + fveg * 4.0f * CIR * pow3i(TV) * DTV;            // func:2854-2857
    SHC = SHC + fveg * CSH * DTV;
    EVC = EVC + fveg * CEV * DESTV * DTV;
    TR = TR + fveg * CTR * DESTV * DTV;

    TV = TV + DTV;                                              // func:2860

    H = RHOAIR * CPAIR * (TAH - SFCTMP) / RAHC;                 // func:2864-2865
    HG = RHOAIR * CPAIR * (TG - TAH) / RAHG;

    QSFC = (0.622f * EAH) / (SFCPRS - 0.378f * EAH);            // func:2868

    if (LITER == 1) break;                                      // func:2870-2875
    if (iter >= 5 && fabsf(DTV) <= 0.01f && LITER == 0) LITER = 1;
  }
  c.iter_vege = iters_done;

  AIR = -EMG * (1.0f - EMV) * LWDN - EMG * EMV * SB * pow4i(TV);  // func:2881-2885
  CIR = EMG * SB;
  CSH = RHOAIR * CPAIR / RAHG;
  CEV = RHOAIR * CPAIR / (GAMMAG * (RAWG + RSURF));
  CGH = 2.0f * DF(ISNOW + 1) / DZSNSO(ISNOW + 1);

  for (int iter = 1; iter <= NITERG; ++iter) {                  // func:2887-2914 loop2
    T = TDC(TG);
    esat(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.0f) { ESTG = ESATW; DESTG = DSATW; }
    else { ESTG = ESATI; DESTG = DSATI; }

    IRG = CIR * pow4i(TG) + AIR;
    SHG = CSH * (TG - TAH);
    EVG = CEV * (ESTG * RHSUR - EAH);
    GH = CGH * (TG - STC(ISNOW + 1));

    float B = SAG - IRG - SHG - EVG - GH;
    float A = 4.0f * CIR * pow3i(TG) + CSH + CEV * DESTG + CGH;
    DTG = B / A;

    IRG = IRG + 4.0f * CIR * pow3i(TG) * DTG;
    SHG = SHG + CSH * DTG;
    EVG = EVG + CEV * DESTG * DTG;
    GH = GH + CGH * DTG;
    TG = TG + DTG;
  }

  if (cx.opt.opt_stc == 1) {                                    // func:2920-2928
    if (snowh > 0.05f && TG > TFRZ) {
      TG = TFRZ;
      IRG = CIR * pow4i(TG) - EMG * (1.0f - EMV) * LWDN - EMG * EMV * SB * pow4i(TV);
      SHG = CSH * (TG - TAH);
      EVG = CEV * (ESTG * RHSUR - EAH);
      GH = SAG - (IRG + SHG + EVG);
    }
  }

  TAUXV = -RHOAIR * CM * UR * UU;                               // func:2932-2933
  TAUYV = -RHOAIR * CM * UR * VV;

  if (cx.opt.opt_sfc == 1 || cx.opt.opt_sfc == 2) {             // func:2943-2957
    CAH2 = FV * KARMAN / logf((2.0f + Z0H) / Z0H);
    CAH2 = FV * KARMAN / (logf((2.0f + Z0H) / Z0H) - FH2);
    float CQ2V = CAH2;
    if (CAH2 < 1.E-5f) {
      T2MV = TAH;
      Q2V = QSFC;
    } else {
      T2MV = TAH - (SHG + SHC / fveg) / (RHOAIR * CPAIR) * 1.0f / CAH2;
      Q2V = QSFC - ((EVC + TR) / fveg + EVG) / (LATHEAV * RHOAIR) * 1.0f / CQ2V;
    }
  }

  CH = CAH;                                                     // func:2960-2962
  CHLEAF = CVH;
  CHUC = 1.0f / RAHG;
}

// func:2967-3257
void bare_flux(const Ctx& cx, ColumnIO& c, int lutyp, float DT, int ISNOW, float SAG, float LWDN,
               float UR, float UU, float VV, float SFCTMP, float THAIR, float QAIR, float EAIR,
               float RHOAIR, float snowh, const A7& DZSNSO, float ZLVL, float ZPD, float Z0M,
               float EMG, const A7& STC, const A7& DF, float RSURF, float LATHEA, float GAMMA,
               float RHSUR, float& TGB, float& CM, float& CH, float& TAUXB, float& TAUYB, float& IRB,
               float& SHB, float& EVB, float& GHB, float& T2MB, float& QSFC, float PSFC, float SFCPRS,
               float& Q2B, float& EHB2) {
  (void)DT; (void)SFCPRS;
  const int NITERB = 5;                                         // func:3115
  float MPE_ = 1.0E-6f;                                         // func:3123-3129
  float DTG = 0.0f;
  int MOZSGN = 0;
  float H = 0.0f;
  float QFX = 0.0f;
  float FV = 0.1f;
  float MOZ = 0.0f, FM = 0.0f, FH = 0.0f, FM2 = 0.0f, FH2 = 0.0f, CH2 = 0.0f, WSTAR = 0.0f;
  float Z0H = 0.0f, EHB = 0.0f, CSH = 0.0f, CEV = 0.0f, ESTG = 0.0f, DESTG = 0.0f;
  float ESATW, ESATI, DSATW, DSATI;

  float CIR = EMG * SB;                                         // func:3131-3132
  float CGH = 2.0f * DF(ISNOW + 1) / DZSNSO(ISNOW + 1);

  for (int iter = 1; iter <= NITERB; ++iter) {                  // func:3134 loop3
    Z0H = Z0M;                                                  // func:3136-3140

    if (cx.opt.opt_sfc == 1) {                                  // func:3142-3148
      sfcdif1(iter, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, Z0H, UR, MPE_, MOZ, MOZSGN, FM, FH, FM2,
              FH2, CM, CH, FV, CH2, c.flags);
    }
    if (cx.opt.opt_sfc == 2) {                                  // func:3150-3164
      sfcdif2(iter, Z0M, TGB, THAIR, UR, cx.tbl.KK_CZIL, ZLVL, CM, CH, MOZ, WSTAR, FV);
      CH = CH / UR;
      CM = CM / UR;
      if (snowh > 0.0f) {
        CM = fmin2(0.01f, CM);
        CH = fmin2(0.01f, CH);
      }
    }

    float RAMB = fmax2(1.0f, 1.0f / (CM * UR));                 // func:3166-3172
    float RAHB = fmax2(1.0f, 1.0f / (CH * UR));
    float RAWB = RAHB;
    float EMB = 1.0f / RAMB;
    (void)EMB;
    EHB = 1.0f / RAHB;

    float T = TDC(TGB);                                         // func:3176-3184
    esat(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.0f) { ESTG = ESATW; DESTG = DSATW; }
    else { ESTG = ESATI; DESTG = DSATI; }

    CSH = RHOAIR * CPAIR / RAHB;                                // func:3186-3187
    CEV = RHOAIR * CPAIR / GAMMA / (RSURF + RAWB);

    IRB = CIR * pow4i(TGB) - EMG * LWDN;                        // func:3191-3194
    SHB = CSH * (TGB - SFCTMP);
    EVB = CEV * (ESTG * RHSUR - EAIR);
    GHB = CGH * (TGB - STC(ISNOW + 1));

    float B = SAG - IRB - SHB - EVB - GHB;                      // func:3196-3198
    float A = 4.0f * CIR * pow3i(TGB) + CSH + CEV * DESTG + CGH;
    DTG = B / A;

    IRB = IRB + 4.0f * CIR * pow3i(TGB) * DTG;                  // func:3200-3203
    SHB = SHB + CSH * DTG;
    EVB = EVB + CEV * DESTG * DTG;
    GHB = GHB + CGH * DTG;

    TGB = TGB + DTG;                                            // func:3206

    H = CSH * (TGB - SFCTMP);                                   // func:3209

    T = TDC(TGB);                                               // func:3211-3217
    esat(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.0f) ESTG = ESATW; else ESTG = ESATI;
    QSFC = 0.622f * (ESTG * RHSUR) / (PSFC - 0.378f * (ESTG * RHSUR));  // func:3218

    QFX = (QSFC - QAIR) * CEV * GAMMA / CPAIR;                  // func:3220
  }
  (void)QFX;

  if (cx.opt.opt_stc == 1) {                                    // func:3225-3233
    if (snowh > 0.05f && TGB > TFRZ) {
      TGB = TFRZ;
      IRB = CIR * pow4i(TGB) - EMG * LWDN;
      SHB = CSH * (TGB - SFCTMP);
      EVB = CEV * (ESTG * RHSUR - EAIR);
      GHB = SAG - (IRB + SHB + EVB);
    }
  }

  TAUXB = -RHOAIR * CM * UR * UU;                               // func:3236-3237
  TAUYB = -RHOAIR * CM * UR * VV;

  if (cx.opt.opt_sfc == 1 || cx.opt.opt_sfc == 2) {             // func:3241-3253
    EHB2 = FV * KARMAN / logf((2.0f + Z0H) / Z0H);
    EHB2 = FV * KARMAN / (logf((2.0f + Z0H) / Z0H) - FH2);
    float CQ2B = EHB2;
    if (EHB2 < 1.0E-5f) {
      T2MB = TGB;
      Q2B = QSFC;
    } else {
      T2MB = TGB - SHB / (RHOAIR * CPAIR) * 1.0f / EHB2;
      Q2B = QSFC - EVB / (LATHEA * RHOAIR) * (1.0f / CQ2B + RSURF);
    }
    if (lutyp == cx.tbl.ISURBAN) Q2B = QSFC;
  }

  CH = EHB;                                                     // func:3256
}

// func:3260-3350
void ragrb(const Ctx& cx, int iter, float VAI, float RHOAIR, float HG, float TAH, float ZPD,
           float Z0MG, float Z0HG, float HCAN, float UC, float Z0H, float FV, float CWP, int lutyp,
           float MPE_, float TV, float& MOZG, float& FHG, float& RAMG, float& RAHG, float& RAWG,
           float& RB) {
  (void)TV;
  MOZG = 0.0f;                                                  // func:3309-3310
  float MOLG = 0.0f;
  if (iter > 1) {                                               // func:3312-3317
    float TMP1 = KARMAN * (GRAV / TAH) * HG / (RHOAIR * CPAIR);
    if (fabsf(TMP1) <= MPE_) TMP1 = MPE_;
    MOLG = -1.0f * pow3i(FV) / TMP1;
    MOZG = fmin2((ZPD - Z0MG) / MOLG, 1.0f);
  }
  float FHGNEW;
  if (MOZG < 0.0f) FHGNEW = powf(1.0f - 15.0f * MOZG, -0.25f);  // func:3319-3323
  else FHGNEW = 1.0f + 4.7f * MOZG;

  if (iter == 1) FHG = FHGNEW;                                  // func:3325-3329
  else FHG = 0.5f * (FHG + FHGNEW);

  float CWPC = powf(CWP * VAI * HCAN * FHG, 0.5f);              // func:3331

  float TMP1 = expf(-CWPC * Z0HG / HCAN);                       // func:3334-3336
  float TMP2 = expf(-CWPC * (Z0H + ZPD) / HCAN);
  float TMPRAH2 = HCAN * expf(CWPC) / CWPC * (TMP1 - TMP2);

  float KH = fmax2(KARMAN * FV * (HCAN - ZPD), MPE_);           // func:3339-3342
  RAMG = 0.0f;
  RAHG = TMPRAH2 / KH;
  RAWG = RAHG;

  float TMPRB = CWPC * 50.0f / (1.0f - expf(-CWPC / 2.0f));     // func:3346-3347
  RB = TMPRB * sqrtf(cx.tbl.LK_DLEAF[lutyp] / UC);
}

// func:3353-3508
void sfcdif1(int iter, float SFCTMP, float RHOAIR, float H, float QAIR, float ZLVL, float ZPD,
             float Z0M, float Z0H, float UR, float MPE_, float& MOZ, int& MOZSGN, float& FM, float& FH,
             float& FM2, float& FH2, float& CM, float& CH, float& FV, float& CH2, uint32_t& flags) {
  float MOZOLD = MOZ;                                           // func:3410
  if (ZLVL <= ZPD) flags |= FLAG_ZLVL;                          // func:3412-3415

  float TMPCM = logf((ZLVL - ZPD) / Z0M);                       // func:3417-3420
  float TMPCH = logf((ZLVL - ZPD) / Z0H);
  float TMPCM2 = logf((2.0f + Z0M) / Z0M);
  float TMPCH2 = logf((2.0f + Z0H) / Z0H);

  float MOL, MOZ2;
  if (iter == 1) {                                              // func:3422-3434
    FV = 0.0f;
    MOZ = 0.0f;
    MOL = 0.0f;
    MOZ2 = 0.0f;
  } else {
    float TVIR = (1.0f + 0.61f * QAIR) * SFCTMP;
    float TMP1 = KARMAN * (GRAV / TVIR) * H / (RHOAIR * CPAIR);
    if (fabsf(TMP1) <= MPE_) TMP1 = MPE_;
    MOL = -1.0f * pow3i(FV) / TMP1;
    MOZ = fmin2((ZLVL - ZPD) / MOL, 1.0f);
    MOZ2 = fmin2((2.0f + Z0H) / MOL, 1.0f);
  }

  if (MOZOLD * MOZ < 0.0f) MOZSGN = MOZSGN + 1;                 // func:3438-3446
  if (MOZSGN >= 2) {
    MOZ = 0.0f;
    FM = 0.0f;
    FH = 0.0f;
    MOZ2 = 0.0f;
    FM2 = 0.0f;
    FH2 = 0.0f;
  }

  float FMNEW, FHNEW, FM2NEW, FH2NEW;
  if (MOZ < 0.0f) {                                             // func:3449-3467
    float TMP1 = powf(1.0f - 16.0f * MOZ, 0.25f);
    float TMP2 = logf((1.0f + TMP1 * TMP1) / 2.0f);
    float TMP3 = logf((1.0f + TMP1) / 2.0f);
    FMNEW = 2.0f * TMP3 + TMP2 - 2.0f * atanf(TMP1) + 1.5707963f;
    FHNEW = 2.0f * TMP2;

    float TMP12 = powf(1.0f - 16.0f * MOZ2, 0.25f);
    float TMP22 = logf((1.0f + TMP12 * TMP12) / 2.0f);
    float TMP32 = logf((1.0f + TMP12) / 2.0f);
    FM2NEW = 2.0f * TMP32 + TMP22 - 2.0f * atanf(TMP12) + 1.5707963f;
    FH2NEW = 2.0f * TMP22;
  } else {
    FMNEW = -5.0f * MOZ;
    FHNEW = FMNEW;
    FM2NEW = -5.0f * MOZ2;
    FH2NEW = FM2NEW;
  }

  if (iter == 1) {                                              // func:3472-3482
    FM = FMNEW;
    FH = FHNEW;
    FM2 = FM2NEW;
    FH2 = FH2NEW;
  } else {
    FM = 0.5f * (FM + FMNEW);
    FH = 0.5f * (FH + FHNEW);
    FM2 = 0.5f * (FM2 + FM2NEW);
    FH2 = 0.5f * (FH2 + FH2NEW);
  }

  FH = fmin2(FH, 0.9f * TMPCH);                                 // func:3486-3489
  FM = fmin2(FM, 0.9f * TMPCM);
  FH2 = fmin2(FH2, 0.9f * TMPCH2);
  FM2 = fmin2(FM2, 0.9f * TMPCM2);

  float CMFM = TMPCM - FM;                                      // func:3491-3501
  float CHFH = TMPCH - FH;
  float CM2FM2 = TMPCM2 - FM2;
  float CH2FH2 = TMPCH2 - FH2;
  if (fabsf(CMFM) <= MPE_) CMFM = MPE_;
  if (fabsf(CHFH) <= MPE_) CHFH = MPE_;
  if (fabsf(CM2FM2) <= MPE_) CM2FM2 = MPE_;
  if (fabsf(CH2FH2) <= MPE_) CH2FH2 = MPE_;
  CM = KARMAN * KARMAN / (CMFM * CMFM);
  CH = KARMAN * KARMAN / (CMFM * CHFH);
  CH2 = KARMAN * KARMAN / (CM2FM2 * CH2FH2);

  FV = UR * sqrtf(CM);                                          // func:3505-3506
  CH2 = KARMAN * FV / CH2FH2;
}

// func:3511-3689
void sfcdif2(int iter, float Z0, float THZ0, float THLM, float SFCSPD, float CZIL, float ZLM,
             float& AKMS, float& AKHS, float& RLMO, float& WSTAR2, float& USTAR) {
  const float WWST = 1.2f;
  const float WWST2 = WWST * WWST;
  const float VKRM = 0.40f;
  const float EXCM = 0.001f;
  const float BETA = 1.0f / 270.0f;
  const float BTG = BETA * GRAV;
  const float ELFC = VKRM * BTG;
  const float WOLD = 0.15f;
  const float WNEW = 1.0f - WOLD;
  const float PIHF = 3.14159265f / 2.0f;
  const float EPSU2 = 1.E-4f;
  const float EPSUST = 0.07f;
  const float ZTMIN = -5.0f;
  const float ZTMAX = 1.0f;
  const float HPBL = 1000.0f;
  const float SQVISC = 258.2f;
  const float RIC = 0.183f;
  const float RRIC = 1.0f / RIC;
  const float FHNEU = 0.8f;
  const float RFC = 0.191f;
  const float RFAC = RIC / (FHNEU * RFC * RFC);

  auto PSLMU = [&](float ZZ) { return -0.96f * logf(1.0f - 4.5f * ZZ); };                      // func:3567
  auto PSLMS = [&](float ZZ) { return ZZ * RRIC - 2.076f * (1.0f - 1.0f / (ZZ + 1.0f)); };     // func:3568
  auto PSLHU = [&](float ZZ) { return -0.96f * logf(1.0f - 4.5f * ZZ); };                      // func:3569
  auto PSLHS = [&](float ZZ) { return ZZ * RFAC - 2.076f * (1.0f - 1.0f / (ZZ + 1.0f)); };     // func:3570
  auto PSPMU = [&](float XX) {
    return -2.0f * logf((XX + 1.0f) * 0.5f) - logf((XX * XX + 1.0f) * 0.5f) + 2.0f * atanf(XX) - PIHF;
  };                                                                                           // func:3572-3574
  auto PSPMS = [&](float YY) { return 5.0f * YY; };                                            // func:3575
  auto PSPHU = [&](float XX) { return -2.0f * logf((XX * XX + 1.0f) * 0.5f); };                // func:3576
  auto PSPHS = [&](float YY) { return 5.0f * YY; };                                            // func:3577

  int ILECH = 0;                                                // func:3586
  float ZILFC = -CZIL * VKRM * SQVISC;                          // func:3589-3593
  float ZU = Z0;
  float RDZ = 1.0f / ZLM;
  float CXCH = EXCM * RDZ;
  float DTHV = THLM - THZ0;

  float DU2 = fmax2(SFCSPD * SFCSPD, EPSU2);                    // func:3596-3597
  float BTGH = BTG * HPBL;

  if (iter == 1) {                                              // func:3599-3607
    if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * powf(fabsf(BTGH * AKHS * DTHV), 2.0f / 3.0f);
    else WSTAR2 = 0.0f;
    USTAR = fmax2(sqrtf(AKMS * sqrtf(DU2 + WSTAR2)), EPSUST);
    RLMO = ELFC * AKHS * DTHV / pow3i(USTAR);
  }

  float ZT = fmax2(1.0E-6f, expf(ZILFC * sqrtf(USTAR * Z0)) * Z0);  // func:3610-3614
  float ZSLU = ZLM + ZU;
  float ZSLT = ZLM + ZT;
  float RLOGU = logf(ZSLU / ZU);
  float RLOGT = logf(ZSLT / ZT);

  float ZETALT = fmax2(ZSLT * RLMO, ZTMIN);                     // func:3617-3621
  RLMO = ZETALT / ZSLT;
  float ZETALU = ZSLU * RLMO;
  float ZETAU = ZU * RLMO;
  float ZETAT = ZT * RLMO;

  float PSMZ, SIMM, PSHZ, SIMH;
  if (ILECH == 0) {                                             // func:3623-3661
    if (RLMO < 0.0f) {
      float XLU4 = 1.0f - 16.0f * ZETALU;
      float XLT4 = 1.0f - 16.0f * ZETALT;
      float XU4 = 1.0f - 16.0f * ZETAU;
      float XT4 = 1.0f - 16.0f * ZETAT;
      float XLU = sqrtf(sqrtf(XLU4));
      float XLT = sqrtf(sqrtf(XLT4));
      float XU = sqrtf(sqrtf(XU4));
      float XT = sqrtf(sqrtf(XT4));
      PSMZ = PSPMU(XU);
      SIMM = PSPMU(XLU) - PSMZ + RLOGU;
      PSHZ = PSPHU(XT);
      SIMH = PSPHU(XLT) - PSHZ + RLOGT;
    } else {
      ZETALU = fmin2(ZETALU, ZTMAX);
      ZETALT = fmin2(ZETALT, ZTMAX);
      PSMZ = PSPMS(ZETAU);
      SIMM = PSPMS(ZETALU) - PSMZ + RLOGU;
      PSHZ = PSPHS(ZETAT);
      SIMH = PSPHS(ZETALT) - PSHZ + RLOGT;
    }
  } else {
    if (RLMO < 0.0f) {
      PSMZ = PSLMU(ZETAU);
      SIMM = PSLMU(ZETALU) - PSMZ + RLOGU;
      PSHZ = PSLHU(ZETAT);
      SIMH = PSLHU(ZETALT) - PSHZ + RLOGT;
    } else {
      ZETALU = fmin2(ZETALU, ZTMAX);
      ZETALT = fmin2(ZETALT, ZTMAX);
      PSMZ = PSLMS(ZETAU);
      SIMM = PSLMS(ZETALU) - PSMZ + RLOGU;
      PSHZ = PSLHS(ZETAT);
      SIMH = PSLHS(ZETALT) - PSHZ + RLOGT;
    }
  }

  USTAR = fmax2(sqrtf(AKMS * sqrtf(DU2 + WSTAR2)), EPSUST);     // func:3664

  ZT = fmax2(1.E-6f, expf(ZILFC * sqrtf(USTAR * Z0)) * Z0);     // func:3667-3676
  ZSLT = ZLM + ZT;
  RLOGT = logf(ZSLT / ZT);
  (void)RLOGT;
  float USTARK = USTAR * VKRM;
  AKMS = fmax2(USTARK / SIMM, CXCH);
  AKHS = fmax2(USTARK / SIMH, CXCH);

  if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * powf(fabsf(BTGH * AKHS * DTHV), 2.0f / 3.0f);  // func:3678-3682
  else WSTAR2 = 0.0f;

  float RLMN = ELFC * AKHS * DTHV / pow3i(USTAR);               // func:3684-3687
  float RLMA = RLMO * WOLD + RLMN * WNEW;
  RLMO = RLMA;
}

// func:3739-3887 (bisection on ci; internal ci2ci at func:3847-3886)
void stomata(const Ctx& cx, int lutyp, float igs, float sfcprs, float sfctmp, float apar, float tv,
             float ea, float ei, float o2, float co2, float foln, float btran, float rb, float& rs,
             float& psn, int& niter_used) {
  const Tables& t = cx.tbl;
  const float CIERR = 5.0E-2f;                                  // func:3782
  const int NITER = 20;                                         // func:3786
  niter_used = 0;

  float cf = sfcprs / (RGAS * sfctmp) * 1.0e06f;                // func:3801-3804
  rs = 1.0f / t.LK_BP[lutyp] * cf;
  psn = 0.0f;
  float ci = co2;
  (void)ci;

  if (apar <= 0.0f) return;                                     // func:3806

  float fnf = fmin2(foln / fmax2(MPE, t.LK_FOLNMX[lutyp]), 1.0f);  // func:3808-3811
  float tc = tv - TFRZ;
  float ppf = 4.6f * apar;
  float j = ppf * t.LK_QE25[lutyp];
  float kc = t.LK_KC25[lutyp] * powf(t.LK_AKC[lutyp], (tc - 25.0f) / 10.0f);  // func:3815-3818
  float ko = t.LK_KO25[lutyp] * powf(t.LK_AKO[lutyp], (tc - 25.0f) / 10.0f);
  float awc = kc * (1.0f + o2 / ko);
  float cp = 0.5f * kc / ko * o2 * 0.21f;
  float vcmx = t.LK_VCMX25[lutyp] /
               (1.0f + expf((-2.2E05f + 710.0f * (tc + TFRZ)) / (8.314f * (tc + TFRZ)))) * fnf * btran *
               (powf(t.LK_AVCMX[lutyp], (tc - 25.0f) / 10.0f));  // func:3821-3824

  float rlb = rb / cf;                                          // func:3827

  float wc = NAN, wj = NAN, we = NAN;                           // func:3854-3856 (SAVEd, NaN-initialised)
  auto ci2ci = [&](float ci_, float& fci_) {                    // func:3847-3886
    if (t.LK_C3C4[lutyp] == 1) {
      wj = fmax2(ci_ - cp, 0.0f) * j / (ci_ + 2.0f * cp);
      wc = fmax2(ci_ - cp, 0.0f) * vcmx / (ci_ + awc);
      we = 0.5f * vcmx;
    } else if (t.LK_C3C4[lutyp] == 2) {
      wj = j;
      wc = vcmx;
      we = 4000.0f * vcmx * ci_ / sfcprs;
    }
    psn = fmin2(fmin2(wj, wc), we) * igs;                       // func:3870
    float cs = fmax2(co2 - 1.37f * rlb * sfcprs * psn, MPE);    // func:3872
    float a = t.LK_MP[lutyp] * psn * sfcprs * ea / (cs * ei) + t.LK_BP[lutyp];
    float b = (t.LK_MP[lutyp] * psn * sfcprs / cs + t.LK_BP[lutyp]) * rlb - 1.0f;
    float c_ = -rlb;
    float q;
    if (b >= 0.0f) q = -0.5f * (b + sqrtf(b * b - 4.0f * a * c_));
    else q = -0.5f * (b - sqrtf(b * b - 4.0f * a * c_));
    float r1 = q / a;
    float r2 = c_ / q;
    rs = fmax2(r1, r2);                                         // func:3883
    fci_ = fmax2(cs - psn * sfcprs * 1.65f * rs, 0.0f);         // func:3885
  };

  float cihigh = 1.5f * co2;                                    // func:3830-3842
  float cilow = 0.0f;
  for (int iter = 1; iter <= NITER; ++iter) {
    niter_used = iter;
    ci = 0.5f * (cihigh + cilow);
    float fci;
    ci2ci(ci, fci);
    if (((cihigh - cilow) <= CIERR) || fabsf(fci - ci) <= MPE) {
      break;
    } else if (fci > ci) {
      cilow = ci;
    } else {
      cihigh = ci;
    }
  }

  rs = rs * cf;                                                 // func:3844
}

// func:3958-3984
void calhum(float SFCTMP, float SFCPRS, float& Q2SAT, float& DQSDT2) {
  const float A2_ = 17.67f, A3_ = 273.15f, A4_ = 29.65f, ELWV = 2.501E6f, A23M4 = A2_ * (A3_ - A4_),
              E0 = 0.611f, RV = 461.0f, EPSILON = 0.622f;
  float ES = E0 * expf(ELWV / RV * (1.0f / A3_ - 1.0f / SFCTMP));  // func:3970
  float SFCPRSX = SFCPRS * 1.E-3f;
  Q2SAT = EPSILON * ES / (SFCPRSX - ES);
  Q2SAT = Q2SAT * 1.E3f;
  DQSDT2 = (Q2SAT / (1.0f + Q2SAT)) * A23M4 / pow2i(SFCTMP - A4_);  // func:3979
  Q2SAT = Q2SAT / 1.E3f;
}

// func:3890-3955
void canres(const Ctx& cx, int lutyp, float sfcprs, float tv, float par, float eah, float btran,
            float& rs, float& psn) {
  const Tables& t = cx.tbl;
  float rcs = 1.0f, rct = 1.0f, rcq = 1.0f;                     // func:3930-3932
  float q2 = 0.622f * eah / (sfcprs - 0.378f * eah);            // func:3935-3937
  q2 = q2 / (1.0f + q2);
  float q2sat, dqsdt2;
  calhum(tv, sfcprs, q2sat, dqsdt2);
  float ff = 2.0f * par / t.LK_RGL[lutyp];                      // func:3940-3942
  rcs = (ff + t.LK_RSMIN[lutyp] / t.LK_RSMAX[lutyp]) / (1.0f + ff);
  rcs = fmin2(fmax2(rcs, 0.0001f), 1.0f);
  rct = 1.0f - 0.0016f * (pow2i(t.LK_TOPT[lutyp] - tv));        // func:3945-3946
  rct = fmin2(fmax2(rct, 0.0001f), 1.0f);
  rcq = 1.0f / (1.0f + t.LK_HS[lutyp] * fmax2(0.0f, q2sat - q2));  // func:3949-3950
  rcq = fmin2(fmax2(rcq, 0.01f), 1.0f);
  rs = t.LK_RSMIN[lutyp] / (rcs * rct * rcq * btran);           // func:3953
  psn = NAN;                                                    // func:3954
}

}  // namespace noahmp_oracle
