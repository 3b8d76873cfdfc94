// TEST INFRASTRUCTURE ONLY (see common.h).
// soilh2o, zwteq, infil, srt, sstep, wdfcnd1, wdfcnd2, groundwater :
// core/module_noahmp_func.f90:5822-6639
#include "physics.h"

namespace noahmp_oracle {

// func:5822-6048
void soilh2o(const Ctx& cx, int sltyp, int lutyp, float DT, const A4& ZSOIL, const A7& DZSNSO,
             int slptyp, float qinsrf, float QSEVA, const A4& ETRANI, const A4& soilice, A4& soilwat,
             A4& SMC, float& ZWT, float& runsrf, float& QDRAIN, float& runsub, A4& WCND,
             float& FCRMAX, int& niter_out) {
  const Tables& t = cx.tbl;
  const float A = 4.0f;                                         // func:5886
  A4 RHSTT, AI, BI, CI, MLIQ, FCR;
  for (int i = 0; i < 4; ++i) { RHSTT.v[i] = 0; AI.v[i] = 0; BI.v[i] = 0; CI.v[i] = 0; MLIQ.v[i] = 0; FCR.v[i] = 0; }
  float FSAT, FFF, RSBMX;

  runsrf = 0.0f;                                                // func:5888-5890
  float qinfil = 0.0f;
  float RSAT = 0.0f;

  for (int iz = 1; iz <= NSOIL; ++iz) {                         // func:5893-5897
    float epore = fmax2(1.0E-4f, (t.LK_SMCMAX[sltyp] - soilice(iz)));
    RSAT = RSAT + fmax2(0.0f, soilwat(iz) - epore) * DZSNSO(iz);
    soilwat(iz) = fmin2(epore, soilwat(iz));
  }

  for (int iz = 1; iz <= NSOIL; ++iz) {                         // func:5900-5904
    float FICE = fmin2(1.0f, soilice(iz) / t.LK_SMCMAX[sltyp]);
    FCR(iz) = fmax2(0.0f, expf(-A * (1.0f - FICE)) - expf(-A)) / (1.0f - expf(-A));
  }

  float SICEMAX = 0.0f;                                         // func:5907-5914
  FCRMAX = 0.0f;
  float SH2OMIN = t.LK_SMCMAX[sltyp];
  for (int iz = 1; iz <= NSOIL; ++iz) {
    if (soilice(iz) > SICEMAX) SICEMAX = soilice(iz);
    if (FCR(iz) > FCRMAX) FCRMAX = FCR(iz);
    if (soilwat(iz) < SH2OMIN) SH2OMIN = soilwat(iz);
  }

  if (cx.opt.opt_run == 2) {                                    // func:5917-5922
    FFF = 2.0f;
    RSBMX = 4.0f;
    zwteq(cx, sltyp, ZSOIL, DZSNSO, soilwat, ZWT);
    runsub = (1.0f - FCRMAX) * RSBMX * expf(-t.KK_TIMEAN) * expf(-FFF * ZWT);
  }

  if (lutyp == t.ISURBAN) FCR(1) = 0.95f;                       // func:5927

  if (cx.opt.opt_run == 1) {                                    // func:5929-5936
    FFF = 6.0f;
    FSAT = t.KK_FSATMAX * expf(-0.5f * FFF * (ZWT - 2.0f));
    if (qinsrf > 0.0f) {
      runsrf = qinsrf * ((1.0f - FCR(1)) * FSAT + FCR(1));
      qinfil = qinsrf - runsrf;
    }
  }

  if (cx.opt.opt_run == 2) {                                    // func:5938-5945
    FFF = 2.0f;
    FSAT = t.KK_FSATMAX * expf(-0.5f * FFF * ZWT);
    if (qinsrf > 0.0f) {
      runsrf = qinsrf * ((1.0f - FCR(1)) * FSAT + FCR(1));
      qinfil = qinsrf - runsrf;
    }
  }

  if (cx.opt.opt_run == 3) {                                    // func:5947-5951
    infil(cx, sltyp, DT, ZSOIL, soilwat, soilice, SICEMAX, qinsrf, qinfil, runsrf);
  }

  if (cx.opt.opt_run == 4) {                                    // func:5953-5968
    float SMCTOT = 0.0f;
    float DZTOT = 0.0f;
    for (int iz = 1; iz <= NSOIL; ++iz) {
      DZTOT = DZTOT + DZSNSO(iz);
      SMCTOT = SMCTOT + SMC(iz) * DZSNSO(iz);
      if (DZTOT >= 2.0f) break;
    }
    SMCTOT = SMCTOT / DZTOT;
    FSAT = powf(fmax2(0.01f, SMCTOT / t.LK_SMCMAX[sltyp]), 4.0f);
    if (qinsrf > 0.0f) {
      runsrf = qinsrf * ((1.0f - FCR(1)) * FSAT + FCR(1));
      qinfil = qinsrf - runsrf;
    }
  }

  int niter = 1;                                                // func:5971-5978
  if (cx.opt.opt_inf == 1) {
    niter = 3;
    if (qinfil * DT > DZSNSO(1) * t.LK_SMCMAX[sltyp]) niter = niter * 2;
  }
  float dtfine = DT / (float)niter;
  niter_out = niter;

  float QDRAIN_SAVE = 0.0f;                                     // func:5981-5996
  for (int iter = 1; iter <= niter; ++iter) {
    srt(cx, sltyp, dtfine, ZSOIL, slptyp, qinfil, ETRANI, QSEVA, soilwat, SMC, ZWT, FCR, SICEMAX,
        FCRMAX, RHSTT, AI, BI, CI, QDRAIN, WCND);
    float WPLUS;
    sstep(cx, sltyp, dtfine, DZSNSO, soilice, soilwat, SMC, AI, BI, CI, RHSTT, WPLUS);
    RSAT = RSAT + WPLUS;
    QDRAIN_SAVE = QDRAIN_SAVE + QDRAIN;
  }

  QDRAIN = QDRAIN_SAVE / (float)niter;                          // func:5998

  runsrf = runsrf * 1000.0f + RSAT * 1000.0f / DT;              // func:6000-6001
  QDRAIN = QDRAIN * 1000.0f;

  if (cx.opt.opt_run == 2) {                                    // func:6004-6014
    float WTSUB = 0.0f;
    for (int iz = 1; iz <= NSOIL; ++iz) WTSUB = WTSUB + WCND(iz) * DZSNSO(iz);
    for (int iz = 1; iz <= NSOIL; ++iz) {
      float MH2O = runsub * DT * (WCND(iz) * DZSNSO(iz)) / WTSUB;
      soilwat(iz) = soilwat(iz) - MH2O / (DZSNSO(iz) * 1000.0f);
    }
  }

  if (cx.opt.opt_run != 1) {                                    // func:6018-6046
    for (int iz = 1; iz <= NSOIL; ++iz) MLIQ(iz) = soilwat(iz) * DZSNSO(iz) * 1000.0f;
    float WATMIN = 0.01f;
    float XS;
    for (int iz = 1; iz <= NSOIL - 1; ++iz) {
      if (MLIQ(iz) < 0.0f) XS = WATMIN - MLIQ(iz);
      else XS = 0.0f;
      MLIQ(iz) = MLIQ(iz) + XS;
      MLIQ(iz + 1) = MLIQ(iz + 1) - XS;
    }
    int iz = NSOIL;
    if (MLIQ(iz) < WATMIN) XS = WATMIN - MLIQ(iz);
    else XS = 0.0f;
    MLIQ(iz) = MLIQ(iz) + XS;
    runsub = runsub - XS / DT;
    for (iz = 1; iz <= NSOIL; ++iz) soilwat(iz) = MLIQ(iz) / (DZSNSO(iz) * 1000.0f);
  }
}

// func:6051-6100
void zwteq(const Ctx& cx, int sltyp, const A4& ZSOIL, const A7& DZSNSO, const A4& soilwat, float& ZWT) {
  const Tables& t = cx.tbl;
  const int NFINE = 100;
  float ZFINE[NFINE + 1];
  float WD1 = 0.0f;                                             // func:6079-6082
  for (int K = 1; K <= NSOIL; ++K) WD1 = WD1 + (t.LK_SMCMAX[sltyp] - soilwat(K)) * DZSNSO(K);
  float DZFINE = 3.0f * (-ZSOIL(NSOIL)) / (float)NFINE;         // func:6084
  for (int K = 1; K <= NFINE; ++K) ZFINE[K] = (float)K * DZFINE;
  ZWT = -3.0f * ZSOIL(NSOIL) - 0.001f;                          // func:6089
  float WD2 = 0.0f;
  for (int K = 1; K <= NFINE; ++K) {                            // func:6092-6099
    float TEMP = 1.0f + (ZWT - ZFINE[K]) / t.LK_PSISAT[sltyp];
    WD2 = WD2 + t.LK_SMCMAX[sltyp] * (1.0f - powf(TEMP, -1.0f / t.LK_BEXP[sltyp])) * DZFINE;
    if (fabsf(WD2 - WD1) <= 0.01f) {
      ZWT = ZFINE[K];
      break;
    }
  }
}

// func:6103-6196
void infil(const Ctx& cx, int sltyp, float DT, const A4& ZSOIL, const A4& soilwat, const A4& soilice,
           float SICEMAX, float qinsrf, float& qinfil, float& runsrf) {
  const Tables& t = cx.tbl;
  const int CVFRZ = 3;                                          // func:6140
  A4 DMAX;
  if (qinsrf > 0.0f) {                                          // func:6142
    float DT1 = DT / 86400.0f;
    float SMCAV = t.LK_SMCMAX[sltyp] - t.LK_SMCWLT[sltyp];
    DMAX(1) = -ZSOIL(1) * SMCAV;                                // func:6147-6151
    float DICE = -ZSOIL(1) * soilice(1);
    DMAX(1) = DMAX(1) * (1.0f - (soilwat(1) + soilice(1) - t.LK_SMCWLT[sltyp]) / SMCAV);
    float DD = DMAX(1);
    for (int K = 2; K <= NSOIL; ++K) {                          // func:6153-6158
      DICE = DICE + (ZSOIL(K - 1) - ZSOIL(K)) * soilice(K);
      DMAX(K) = (ZSOIL(K - 1) - ZSOIL(K)) * SMCAV;
      DMAX(K) = DMAX(K) * (1.0f - (soilwat(K) + soilice(K) - t.LK_SMCWLT[sltyp]) / SMCAV);
      DD = DD + DMAX(K);
    }
    float VAL = (1.0f - expf(-t.LK_KDT[sltyp] * DT1));          // func:6160-6163
    float DDT = DD * VAL;
    float PX = fmax2(0.0f, qinsrf * DT);
    float INFMAX = (PX * (DDT / (PX + DDT))) / DT;

    float FCR = 1.0f;                                           // func:6167-6180
    if (DICE > 1.0E-2f) {
      float ACRT = (float)CVFRZ * t.LK_FRZX[sltyp] / DICE;
      float SUM = 1.0f;
      int IALP1 = CVFRZ - 1;
      for (int J = 1; J <= IALP1; ++J) {
        int K = 1;
        for (int JJ = J + 1; JJ <= IALP1; ++JJ) K = K * JJ;
        // ACRT ** (CVFRZ - J): integer exponent 2 then 1
        float p = (CVFRZ - J == 2) ? ACRT * ACRT : ACRT;
        SUM = SUM + p / (float)K;
      }
      FCR = 1.0f - expf(-ACRT) * SUM;
    }
    INFMAX = INFMAX * FCR;                                      // func:6184

    float WDF, WCND;
    wdfcnd2(cx, sltyp, WDF, WCND, soilwat(1), SICEMAX);         // func:6189-6191
    INFMAX = fmax2(INFMAX, WCND);
    INFMAX = fmin2(INFMAX, PX);

    runsrf = fmax2(0.0f, qinsrf - INFMAX);                      // func:6193-6194
    qinfil = qinsrf - runsrf;
  }
}

// func:6199-6305
void srt(const Ctx& cx, int sltyp, float DT, const A4& ZSOIL, int slptyp, float qinfil,
         const A4& ETRANI, float QSEVA, const A4& soilwat, const A4& SMC, float ZWT, const A4& FCR,
         float SICEMAX, float FCRMAX, A4& RHSTT, A4& AI, A4& BI, A4& CI, float& QDRAIN, A4& WCND) {
  (void)DT; (void)ZWT;
  const Tables& t = cx.tbl;
  A4 DDZ, DENOM, DSMDZ, WFLUX, WDF, SMX;
  for (int i = 0; i < 4; ++i) { DDZ.v[i] = 0; DENOM.v[i] = 0; DSMDZ.v[i] = 0; WFLUX.v[i] = 0; WDF.v[i] = 0; SMX.v[i] = 0; }
  float TEMP1;

  if (cx.opt.opt_inf == 1) {                                    // func:6246-6251
    for (int K = 1; K <= NSOIL; ++K) {
      wdfcnd1(cx, sltyp, WDF(K), WCND(K), SMC(K), FCR(K));
      SMX(K) = SMC(K);
    }
  }
  if (cx.opt.opt_inf == 2) {                                    // func:6253-6258
    for (int K = 1; K <= NSOIL; ++K) {
      wdfcnd2(cx, sltyp, WDF(K), WCND(K), soilwat(K), SICEMAX);
      SMX(K) = soilwat(K);
    }
  }

  for (int K = 1; K <= NSOIL; ++K) {                            // func:6260-6287
    if (K == 1) {
      DENOM(K) = -ZSOIL(K);
      TEMP1 = -ZSOIL(K + 1);
      DDZ(K) = 2.0f / TEMP1;
      DSMDZ(K) = 2.0f * (SMX(K) - SMX(K + 1)) / TEMP1;
      WFLUX(K) = WDF(K) * DSMDZ(K) + WCND(K) - qinfil + ETRANI(K) + QSEVA;
    } else if (K < NSOIL) {
      DENOM(K) = (ZSOIL(K - 1) - ZSOIL(K));
      TEMP1 = (ZSOIL(K - 1) - ZSOIL(K + 1));
      DDZ(K) = 2.0f / TEMP1;
      DSMDZ(K) = 2.0f * (SMX(K) - SMX(K + 1)) / TEMP1;
      WFLUX(K) = WDF(K) * DSMDZ(K) + WCND(K) - WDF(K - 1) * DSMDZ(K - 1) - WCND(K - 1) + ETRANI(K);
    } else {
      DENOM(K) = (ZSOIL(K - 1) - ZSOIL(K));
      if (cx.opt.opt_run == 1 || cx.opt.opt_run == 2) QDRAIN = 0.0f;
      if (cx.opt.opt_run == 3) QDRAIN = t.LK_SLOPE[slptyp] * WCND(K);
      if (cx.opt.opt_run == 4) QDRAIN = (1.0f - FCRMAX) * WCND(K);
      WFLUX(K) = -(WDF(K - 1) * DSMDZ(K - 1)) - WCND(K - 1) + ETRANI(K) + QDRAIN;
    }
  }

  for (int K = 1; K <= NSOIL; ++K) {                            // func:6289-6304
    if (K == 1) {
      AI(K) = 0.0f;
      BI(K) = WDF(K) * DDZ(K) / DENOM(K);
      CI(K) = -BI(K);
    } else if (K < NSOIL) {
      AI(K) = -WDF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = -WDF(K) * DDZ(K) / DENOM(K);
      BI(K) = -(AI(K) + CI(K));
    } else {
      AI(K) = -WDF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = 0.0f;
      BI(K) = -(AI(K) + CI(K));
    }
    RHSTT(K) = WFLUX(K) / (-DENOM(K));
  }
}

// func:6308-6383
void sstep(const Ctx& cx, int sltyp, float DT, const A7& DZSNSO, const A4& soilice, A4& soilwat,
           A4& SMC, A4& AI, A4& BI, A4& CI, A4& RHSTT, float& WPLUS) {
  const Tables& t = cx.tbl;
  A4 RHSTTIN, CIIN;
  WPLUS = 0.0f;                                                 // func:6345
  for (int K = 1; K <= NSOIL; ++K) {                            // func:6347-6352
    RHSTT(K) = RHSTT(K) * DT;
    AI(K) = AI(K) * DT;
    BI(K) = 1.0f + BI(K) * DT;
    CI(K) = CI(K) * DT;
  }
  for (int K = 1; K <= NSOIL; ++K) {                            // func:6356-6359
    RHSTTIN(K) = RHSTT(K);
    CIIN(K) = CI(K);
  }
  rosr12(CI.v, AI.v, BI.v, CIIN.v, RHSTTIN.v, RHSTT.v, 1, NSOIL, 0);  // func:6363

  for (int K = 1; K <= NSOIL; ++K) soilwat(K) = soilwat(K) + CI(K);   // func:6365-6367

  float epore;
  for (int K = NSOIL; K >= 2; --K) {                            // func:6372-6377
    epore = fmax2(1.0E-4f, t.LK_SMCMAX[sltyp] - soilice(K));
    WPLUS = fmax2(soilwat(K) - epore, 0.0f) * DZSNSO(K);
    soilwat(K) = fmin2(epore, soilwat(K));
    soilwat(K - 1) = soilwat(K - 1) + WPLUS / DZSNSO(K - 1);
  }
  epore = fmax2(1.0E-4f, t.LK_SMCMAX[sltyp] - soilice(1));      // func:6379-6382
  WPLUS = fmax2(soilwat(1) - epore, 0.0f) * DZSNSO(1);
  soilwat(1) = fmin2(epore, soilwat(1));
  for (int K = 1; K <= NSOIL; ++K) SMC(K) = soilwat(K) + soilice(K);
}

// func:6386-6417
void wdfcnd1(const Ctx& cx, int sltyp, float& WDF, float& WCND, float SMC, float FCR) {
  const Tables& t = cx.tbl;
  float FACTR = fmax2(0.01f, SMC / t.LK_SMCMAX[sltyp]);
  float EXPON = t.LK_BEXP[sltyp] + 2.0f;
  WDF = t.LK_DWSAT[sltyp] * powf(FACTR, EXPON);
  WDF = WDF * (1.0f - FCR);
  EXPON = 2.0f * t.LK_BEXP[sltyp] + 3.0f;
  WCND = t.LK_DKSAT[sltyp] * powf(FACTR, EXPON);
  WCND = WCND * (1.0f - FCR);
}

// func:6420-6455
void wdfcnd2(const Ctx& cx, int sltyp, float& WDF, float& WCND, float SMC, float soilice) {
  const Tables& t = cx.tbl;
  float FACTR = fmax2(0.01f, SMC / t.LK_SMCMAX[sltyp]);
  float EXPON = t.LK_BEXP[sltyp] + 2.0f;
  WDF = t.LK_DWSAT[sltyp] * powf(FACTR, EXPON);
  if (soilice > 0.0f) {
    float VKWGT = 1.0f / (1.0f + powf(500.0f * soilice, 3.0f));
    WDF = VKWGT * WDF + (1.0f - VKWGT) * t.LK_DWSAT[sltyp] * powf(0.2f / t.LK_SMCMAX[sltyp], EXPON);
  }
  EXPON = 2.0f * t.LK_BEXP[sltyp] + 3.0f;
  WCND = t.LK_DKSAT[sltyp] * powf(FACTR, EXPON);
}

// func:6458-6639
void groundwater(const Ctx& cx, int sltyp, float DT, const A4& ZSOIL, const A4& soilice,
                 const A4& WCND, float FCRMAX, A4& soilwat, float& ZWT, float& WA, float& WT,
                 float& QIN, float& QDIS) {
  const Tables& t = cx.tbl;
  const float ROUS = 0.2f, CMIC = 0.20f;                        // func:6511-6512
  A4 DZMM, ZNODE, MLIQ, epore, HK, SMC;
  QDIS = 0.0f;                                                  // func:6515-6516
  QIN = 0.0f;

  DZMM(1) = -ZSOIL(1) * 1.0E3f;                                 // func:6521-6524
  for (int IZ = 2; IZ <= NSOIL; ++IZ) DZMM(IZ) = 1.0E3f * (ZSOIL(IZ - 1) - ZSOIL(IZ));

  ZNODE(1) = -ZSOIL(1) / 2.0f;                                  // func:6528-6531
  for (int IZ = 2; IZ <= NSOIL; ++IZ) ZNODE(IZ) = -ZSOIL(IZ - 1) + 0.5f * (ZSOIL(IZ - 1) - ZSOIL(IZ));

  for (int IZ = 1; IZ <= NSOIL; ++IZ) {                         // func:6535-6540
    SMC(IZ) = soilwat(IZ) + soilice(IZ);
    MLIQ(IZ) = soilwat(IZ) * DZMM(IZ);
    epore(IZ) = fmax2(0.01f, t.LK_SMCMAX[sltyp] - soilice(IZ));
    HK(IZ) = 1.0E3f * WCND(IZ);
  }

  int IWT = NSOIL;                                              // func:6545-6551
  for (int IZ = 2; IZ <= NSOIL; ++IZ) {
    if (ZWT <= -ZSOIL(IZ)) {
      IWT = IZ - 1;
      break;
    }
  }

  float FFF = 6.0f;                                             // func:6554-6557
  float RSBMX = 5.0f;
  QDIS = (1.0f - FCRMAX) * RSBMX * expf(-t.KK_TIMEAN) * expf(-FFF * (ZWT - 2.0f));

  double S_NODE = (double)fmin2(1.0f, SMC(IWT) / t.LK_SMCMAX[sltyp]);  // func:6560 (real(8))
  const double S_MIN = (double)0.01f;                           // func:6561 real(0.01,KIND=8)
  S_NODE = S_NODE > S_MIN ? S_NODE : S_MIN;
  float SMPFZ = (float)(-((double)(t.LK_PSISAT[sltyp] * 1000.0f) * pow(S_NODE, (double)(-t.LK_BEXP[sltyp]))));  // func:6562
  SMPFZ = fmax2(-120000.0f, CMIC * SMPFZ);                      // func:6563

  float KA = HK(IWT);                                           // func:6566

  float WH_ZWT = -ZWT * 1.0E3f;                                 // func:6568-6571
  float WH = SMPFZ - ZNODE(IWT) * 1.0E3f;
  QIN = -KA * (WH_ZWT - WH) / ((ZWT - ZNODE(IWT)) * 1.0E3f);
  QIN = fmax2(-10.0f / DT, fmin2(10.0f / DT, QIN));

  WT = WT + (QIN - QDIS) * DT;                                  // func:6575

  if (IWT == NSOIL) {                                           // func:6577-6584
    WA = WA + (QIN - QDIS) * DT;
    WT = WA;
    ZWT = (-ZSOIL(NSOIL) + 25.0f) - WA / 1000.0f / ROUS;
    MLIQ(NSOIL) = MLIQ(NSOIL) - QIN * DT;
    MLIQ(NSOIL) = MLIQ(NSOIL) + fmax2(0.0f, WA - 5000.0f);
    WA = fmin2(WA, 5000.0f);
  } else {
    if (IWT == NSOIL - 1) {                                     // func:6587-6597
      ZWT = -ZSOIL(NSOIL) - (WT - ROUS * 1000.0f * 25.0f) / (epore(NSOIL)) / 1000.0f;
    } else {
      float WS = 0.0f;
      for (int IZ = IWT + 2; IZ <= NSOIL; ++IZ) WS = WS + epore(IZ) * DZMM(IZ);
      ZWT = -ZSOIL(IWT + 1) - (WT - ROUS * 1000.0f * 25.0f - WS) / (epore(IWT + 1)) / 1000.0f;
    }
    float WTSUB = 0.0f;                                         // func:6599-6606
    for (int IZ = 1; IZ <= NSOIL; ++IZ) WTSUB = WTSUB + HK(IZ) * DZMM(IZ);
    for (int IZ = 1; IZ <= NSOIL; ++IZ) MLIQ(IZ) = MLIQ(IZ) - QDIS * DT * HK(IZ) * DZMM(IZ) / WTSUB;
  }

  ZWT = fmax2(1.5f, ZWT);                                       // func:6609

  float WATMIN = 0.01f;                                         // func:6615-6634
  float XS;
  for (int IZ = 1; IZ <= NSOIL - 1; ++IZ) {
    if (MLIQ(IZ) < 0.0f) XS = WATMIN - MLIQ(IZ);
    else XS = 0.0f;
    MLIQ(IZ) = MLIQ(IZ) + XS;
    MLIQ(IZ + 1) = MLIQ(IZ + 1) - XS;
  }
  int IZ = NSOIL;
  if (MLIQ(IZ) < WATMIN) XS = WATMIN - MLIQ(IZ);
  else XS = 0.0f;
  MLIQ(IZ) = MLIQ(IZ) + XS;
  WA = WA - XS;
  WT = WT - XS;

  for (IZ = 1; IZ <= NSOIL; ++IZ) soilwat(IZ) = MLIQ(IZ) / DZMM(IZ);  // func:6636-6638
}

}  // namespace noahmp_oracle
