// TEST INFRASTRUCTURE ONLY (see common.h).
// radiation, albedo, surrad, snowage, snowalb_*, groundalb, twostream :
// core/module_noahmp_func.f90:1598-2462
#include "physics.h"

namespace noahmp_oracle {

// func:1598-1714
void radiation(const Ctx& cx, int lutyp, int IST, int ISC, int ICE, float SNEQVO, float sneqv,
               float DT, float COSZ, float snowh, float TG, float TV, float FSNO, float QSNOW,
               float FWET, float elai, float esai, const A4& SMC, const A2& SOLAD, const A2& SOLAI,
               float fveg, float& ALBOLD, float& TAUSS, float& FSUN, float& LAISUN, float& LAISHA,
               float& PARSUN, float& PARSHA, float& SAV, float& SAG, float& FSR, float& FSA,
               float& FSRV, float& FSRG, float& BGAP, float& WGAP) {
  A2 ALBGRD, ALBGRI, ALBD, ALBI, FABD, FABI, FTDD, FTID, FTII, FREVI, FREVD, FREGI, FREGD;
  // FREV*/FREG* are left undefined by the Fortran at night (albedo returns at func:1823); zeroed here
  for (int ib = 1; ib <= 2; ++ib) { FREVI(ib) = 0.0f; FREVD(ib) = 0.0f; FREGI(ib) = 0.0f; FREGD(ib) = 0.0f; }
  float FAGE = 0.0f;
  const float MPE_ = 1.0E-6f;                                   // func:1677

  albedo(cx, lutyp, IST, ISC, ICE, DT, COSZ, FAGE, elai, esai, TG, TV, snowh, FSNO, FWET, SMC, SNEQVO,
         sneqv, QSNOW, fveg, ALBOLD, TAUSS, ALBGRD, ALBGRI, ALBD, ALBI, FABD, FABI, FTDD, FTID, FTII,
         FSUN, FREVI, FREVD, FREGD, FREGI, BGAP, WGAP);         // func:1682-1691

  float FSHA = 1.0f - FSUN;                                     // func:1695-1698
  LAISUN = elai * FSUN;
  LAISHA = elai * FSHA;
  float VAI = elai + esai;

  surrad(MPE_, FSUN, FSHA, elai, VAI, LAISUN, LAISHA, SOLAD, SOLAI, FABD, FABI, FTDD, FTID, FTII,
         ALBGRD, ALBGRI, ALBD, ALBI, PARSUN, PARSHA, SAV, SAG, FSA, FSR, FREVI, FREVD, FREGD, FREGI,
         FSRV, FSRG);                                           // func:1705-1712
}

// func:1717-1887
void albedo(const Ctx& cx, int lutyp, int IST, int ISC, int ICE, float DT, float COSZ, float& FAGE,
            float elai, float esai, float TG, float TV, float snowh, float FSNO, float FWET,
            const A4& SMC, float SNEQVO, float sneqv, float QSNOW, float fveg, float& ALBOLD,
            float& TAUSS, A2& ALBGRD, A2& ALBGRI, A2& ALBD, A2& ALBI, A2& FABD, A2& FABI, A2& FTDD,
            A2& FTID, A2& FTII, float& FSUN, A2& FREVI, A2& FREVD, A2& FREGD, A2& FREGI, float& BGAP,
            float& WGAP) {
  (void)snowh;
  const Tables& t = cx.tbl;
  const float MPE_ = 1.0E-06f;                                  // func:1804
  BGAP = 0.0f;
  WGAP = 0.0f;
  for (int IB = 1; IB <= NBAND; ++IB) {                         // func:1810-1821
    ALBD(IB) = 0.0f; ALBI(IB) = 0.0f; ALBGRD(IB) = 0.0f; ALBGRI(IB) = 0.0f;
    FABD(IB) = 0.0f; FABI(IB) = 0.0f; FTDD(IB) = 0.0f; FTID(IB) = 0.0f; FTII(IB) = 0.0f;
    if (IB == 1) FSUN = 0.0f;
  }
  if (COSZ <= 0.0f) return;                                     // func:1823

  A2 RHO, TAU, FTDI, ALBSND, ALBSNI;
  float VAI = 0.0f, WL, WS;
  for (int IB = 1; IB <= NBAND; ++IB) {                         // func:1827-1833
    VAI = elai + esai;
    WL = elai / fmax2(VAI, MPE_);
    WS = esai / fmax2(VAI, MPE_);
    RHO(IB) = fmax2(t.LK_RHOL[IB][lutyp] * WL + t.LK_RHOS[IB][lutyp] * WS, MPE_);
    TAU(IB) = fmax2(t.LK_TAUL[IB][lutyp] * WL + t.LK_TAUS[IB][lutyp] * WS, MPE_);
  }

  snowage(cx, DT, TG, SNEQVO, sneqv, TAUSS, FAGE);              // func:1837

  for (int IB = 1; IB <= NBAND; ++IB) { ALBSND(IB) = 0.0f; ALBSNI(IB) = 0.0f; }
  if (cx.opt.opt_alb == 1) snowalb_bats(FSNO, COSZ, FAGE, ALBSND, ALBSNI);  // func:1841-1842
  if (cx.opt.opt_alb == 2) {                                                // func:1843-1846
    float ALB;
    snowalb_class(cx, QSNOW, DT, ALB, ALBOLD, ALBSND, ALBSNI);
    ALBOLD = ALB;
  }

  groundalb(cx, ICE, IST, ISC, FSNO, SMC, ALBSND, ALBSNI, COSZ, TG, ALBGRD, ALBGRI);  // func:1850-1853

  float GDIR = 0.0f;
  for (int ib = 1; ib <= NBAND; ++ib) {                         // func:1858-1873
    int ic = 0;
    twostream(cx, lutyp, ib, ic, COSZ, VAI, FWET, TV, ALBGRD, ALBGRI, RHO, TAU, fveg, IST, FABD, ALBD,
              FTDD, FTID, GDIR, FREVD, FREGD, BGAP, WGAP);
    ic = 1;
    twostream(cx, lutyp, ib, ic, COSZ, VAI, FWET, TV, ALBGRD, ALBGRI, RHO, TAU, fveg, IST, FABI, ALBI,
              FTDI, FTII, GDIR, FREVI, FREGI, BGAP, WGAP);
  }

  float EXT = GDIR / COSZ * sqrtf(1.0f - RHO(1) - TAU(1));      // func:1877
  FSUN = (1.0f - expf(-EXT * VAI)) / fmax2(EXT * VAI, MPE_);    // func:1878
  EXT = FSUN;
  if (EXT < 0.01f) WL = 0.0f; else WL = EXT;                    // func:1881-1885
  FSUN = WL;
}

// func:1890-2005
void surrad(float MPE_, float FSUN, float FSHA, float elai, float VAI, float LAISUN, float LAISHA,
            const A2& SOLAD, const A2& SOLAI, const A2& FABD, const A2& FABI, const A2& FTDD,
            const A2& FTID, const A2& FTII, const A2& ALBGRD, const A2& ALBGRI, const A2& ALBD,
            const A2& ALBI, float& PARSUN, float& PARSHA, float& SAV, float& SAG, float& FSA,
            float& FSR, const A2& FREVI, const A2& FREVD, const A2& FREGD, const A2& FREGI,
            float& FSRV, float& FSRG) {
  A2 CAD, CAI;
  SAG = 0.0f;                                                   // func:1955-1957
  SAV = 0.0f;
  FSA = 0.0f;
  for (int IB = 1; IB <= NBAND; ++IB) {                         // func:1961-1980
    CAD(IB) = SOLAD(IB) * FABD(IB);
    CAI(IB) = SOLAI(IB) * FABI(IB);
    SAV = SAV + CAD(IB) + CAI(IB);
    FSA = FSA + CAD(IB) + CAI(IB);
    float TRD = SOLAD(IB) * FTDD(IB);
    float TRI = SOLAD(IB) * FTID(IB) + SOLAI(IB) * FTII(IB);
    float ABS = TRD * (1.0f - ALBGRD(IB)) + TRI * (1.0f - ALBGRI(IB));
    SAG = SAG + ABS;
    FSA = FSA + ABS;
  }
  float LAIFRA = elai / fmax2(VAI, MPE_);                       // func:1985
  if (FSUN > 0.0f) {                                            // func:1986-1992
    PARSUN = (CAD(1) + FSUN * CAI(1)) * LAIFRA / fmax2(LAISUN, MPE_);
    PARSHA = (FSHA * CAI(1)) * LAIFRA / fmax2(LAISHA, MPE_);
  } else {
    PARSUN = 0.0f;
    PARSHA = (CAD(1) + CAI(1)) * LAIFRA / fmax2(LAISHA, MPE_);
  }
  float RVIS = ALBD(1) * SOLAD(1) + ALBI(1) * SOLAI(1);         // func:1996-1998
  float RNIR = ALBD(2) * SOLAD(2) + ALBI(2) * SOLAI(2);
  FSR = RVIS + RNIR;
  FSRV = FREVD(1) * SOLAD(1) + FREVI(1) * SOLAI(1) + FREVD(2) * SOLAD(2) + FREVI(2) * SOLAI(2);  // func:2001
  FSRG = FREGD(1) * SOLAD(1) + FREGI(1) * SOLAI(1) + FREGD(2) * SOLAD(2) + FREGI(2) * SOLAI(2);  // func:2002
}

// func:2008-2054
void snowage(const Ctx& cx, float DT, float TG, float SNEQVO, float sneqv, float& TAUSS, float& FAGE) {
  if (sneqv <= 0.0f) {
    TAUSS = 0.0f;
  } else if (sneqv > 800.0f) {
    TAUSS = 0.0f;
  } else {
    float DELA0 = 1.0E-6f * DT;                                 // func:2040
    float ARG = 5.0E3f * (1.0f / TFRZ - 1.0f / TG);
    float AGE1 = expf(ARG);
    float AGE2 = expf(fmin2(0.0f, 10.0f * ARG));
    float AGE3 = 0.3f;
    float TAGE = AGE1 + AGE2 + AGE3;
    float DELA = DELA0 * TAGE;
    float DELS = fmax2(0.0f, sneqv - SNEQVO) / cx.tbl.KK_SWEMAX;
    float SGE = (TAUSS + DELA) * (1.0f - DELS);
    TAUSS = fmax2(0.0f, SGE);                                   // func:2049
  }
  FAGE = TAUSS / (TAUSS + 1.0f);                                // func:2052
}

// func:2057-2102
void snowalb_bats(float FSNO, float COSZ, float FAGE, A2& ALBSND, A2& ALBSNI) {
  (void)FSNO;
  const float C1 = 0.2f, C2 = 0.5f;
  for (int ib = 1; ib <= NBAND; ++ib) { ALBSND(ib) = 0.0f; ALBSNI(ib) = 0.0f; }
  float SL = 2.0f;
  float SL1 = 1.0f / SL;
  float SL2 = 2.0f * SL;
  float CF1 = ((1.0f + SL1) / (1.0f + SL2 * COSZ) - SL1);      // func:2093
  float FZEN = fmax2(CF1, 0.0f);
  ALBSNI(1) = 0.95f * (1.0f - C1 * FAGE);
  ALBSNI(2) = 0.65f * (1.0f - C2 * FAGE);
  ALBSND(1) = ALBSNI(1) + 0.4f * FZEN * (1.0f - ALBSNI(1));
  ALBSND(2) = ALBSNI(2) + 0.4f * FZEN * (1.0f - ALBSNI(2));
}

// func:2105-2151
void snowalb_class(const Ctx& cx, float QSNOW, float DT, float& ALB, float ALBOLD, A2& ALBSND,
                   A2& ALBSNI) {
  for (int ib = 1; ib <= NBAND; ++ib) { ALBSND(ib) = 0.0f; ALBSNI(ib) = 0.0f; }
  ALB = 0.55f + (ALBOLD - 0.55f) * expf(-0.01f * DT / 3600.0f);  // func:2137
  if (QSNOW > 0.0f) {                                            // func:2142-2144
    ALB = ALB + fmin2(QSNOW * DT, cx.tbl.KK_SWEMAX) * (0.84f - ALB) / cx.tbl.KK_SWEMAX;
  }
  ALBSNI(1) = ALB;
  ALBSNI(2) = ALB;
  ALBSND(1) = ALB;
  ALBSND(2) = ALB;
}

// func:2154-2212
void groundalb(const Ctx& cx, int ICE, int IST, int ISC, float FSNO, const A4& SMC, const A2& ALBSND,
               const A2& ALBSNI, float COSZ, float TG, A2& ALBGRD, A2& ALBGRI) {
  (void)ICE;
  const Tables& t = cx.tbl;
  for (int IB = 1; IB <= NBAND; ++IB) {
    float INC = fmax2(0.11f - 0.40f * SMC(1), 0.0f);            // func:2189
    float ALBSOD, ALBSOI;
    if (IST == 1) {
      ALBSOD = fmin2(t.LK_ALBSAT[IB][ISC] + INC, t.LK_ALBDRY[IB][ISC]);  // func:2191
      ALBSOI = ALBSOD;
    } else if (TG > TFRZ) {
      ALBSOD = 0.06f / (powf(fmax2(0.01f, COSZ), 1.7f) + 0.15f);         // func:2194
      ALBSOI = 0.06f;
    } else {
      ALBSOD = t.KK_ALBLAKE[IB];                                          // func:2197
      ALBSOI = ALBSOD;
    }
    if (IST == 1 && ISC == 9) {                                 // func:2203-2206
      ALBSOD = ALBSOD + 0.10f;
      ALBSOI = ALBSOI + 0.10f;
    }
    ALBGRD(IB) = ALBSOD * (1.0f - FSNO) + ALBSND(IB) * FSNO;    // func:2208-2209
    ALBGRI(IB) = ALBSOI * (1.0f - FSNO) + ALBSNI(IB) * FSNO;
  }
}

// func:2215-2462
void twostream(const Ctx& cx, int lutyp, int IB, int IC, float COSZ, float VAI, float FWET, float T,
               const A2& ALBGRD, const A2& ALBGRI, const A2& RHO, const A2& TAU, float fveg, int IST,
               A2& FAB, A2& FRE, A2& FTD, A2& FTI, float& GDIR, A2& FREV, A2& FREG, float& BGAP,
               float& WGAP) {
  (void)IST;
  const Tables& t = cx.tbl;
  const float PAI = 3.14159265f;                                // func:2292
  float GAP = 0.0f, KOPEN = 0.0f;
  if (VAI == 0.0f) {                                            // func:2307-2309
    GAP = 1.0f;
    KOPEN = 1.0f;
  } else {
    if (cx.opt.opt_rad == 1) {                                  // func:2311-2324
      float RC = t.LK_RCROWN[lutyp];
      float DENfveg = -logf(fmax2(1.0f - fveg, 0.01f)) / (PAI * pow2i(RC));
      float HD = t.LK_HVT[lutyp] - t.LK_HVB[lutyp];
      float BB = 0.5f * HD;
      float THETAP = atanf(BB / RC * tanf(acosf(fmax2(0.01f, COSZ))));
      BGAP = expf(-DENfveg * PAI * pow2i(RC) / cosf(THETAP));
      float FA = VAI / (1.33f * PAI * powf(RC, 3.0f) * (BB / RC) * DENfveg);
      float NEWVAI = HD * FA;
      WGAP = (1.0f - BGAP) * expf(-0.5f * NEWVAI / COSZ);
      GAP = fmin2(1.0f - fveg, BGAP + WGAP);
      KOPEN = 0.05f;
    }
    if (cx.opt.opt_rad == 2) {                                  // func:2326-2329
      GAP = 0.0f;
      KOPEN = 0.0f;
    }
    if (cx.opt.opt_rad == 3) {                                  // func:2331-2334
      GAP = 1.0f - fveg;
      KOPEN = 1.0f - fveg;
    }
  }

  float COSZI = fmax2(0.001f, COSZ);                            // func:2344
  float CHIL = fmin2(fmax2(t.LK_XL[lutyp], -0.4f), 0.6f);
  if (fabsf(CHIL) <= 0.01f) CHIL = 0.01f;
  float PHI1 = 0.5f - 0.633f * CHIL - 0.330f * CHIL * CHIL;
  float PHI2 = 0.877f * (1.0f - 2.0f * PHI1);
  GDIR = PHI1 + PHI2 * COSZI;
  float EXT = GDIR / COSZI;
  float AVMU = (1.0f - PHI1 / PHI2 * logf((PHI1 + PHI2) / PHI1)) / PHI2;  // func:2351
  float OMEGAL = RHO(IB) + TAU(IB);
  float TMP0 = GDIR + PHI2 * COSZI;
  float TMP1 = PHI1 * COSZI;
  float ASU = 0.5f * OMEGAL * GDIR / TMP0 * (1.0f - TMP1 / TMP0 * logf((TMP1 + TMP0) / TMP1));  // func:2355
  float BETADL = (1.0f + AVMU * EXT) / (OMEGAL * AVMU * EXT) * ASU;                             // func:2356
  float BETAIL = 0.5f * (RHO(IB) + TAU(IB) + (RHO(IB) - TAU(IB)) * pow2i((1.0f + CHIL) / 2.0f)) /
                 OMEGAL;                                        // func:2357-2358

  float TMP2;
  if (T > TFRZ) {                                               // func:2362-2370
    TMP0 = OMEGAL;
    TMP1 = BETADL;
    TMP2 = BETAIL;
  } else {
    TMP0 = (1.0f - FWET) * OMEGAL + FWET * t.KK_OMEGAS[IB];
    TMP1 = ((1.0f - FWET) * OMEGAL * BETADL + FWET * t.KK_OMEGAS[IB] * t.KK_BETADS) / TMP0;
    TMP2 = ((1.0f - FWET) * OMEGAL * BETAIL + FWET * t.KK_OMEGAS[IB] * t.KK_BETAIS) / TMP0;
  }
  float OMEGA = TMP0;
  float BETAD = TMP1;
  float BETAI = TMP2;

  float B = 1.0f - OMEGA + OMEGA * BETAI;                       // func:2378-2392
  float C = OMEGA * BETAI;
  TMP0 = AVMU * EXT;
  float D = TMP0 * OMEGA * BETAD;
  float F = TMP0 * OMEGA * (1.0f - BETAD);
  TMP1 = B * B - C * C;
  float H = sqrtf(TMP1) / AVMU;
  float SIGMA = TMP0 * TMP0 - TMP1;
  if (fabsf(SIGMA) < 1.e-6f) SIGMA = copysignf(1.e-6f, SIGMA);
  float P1 = B + AVMU * H;
  float P2 = B - AVMU * H;
  float P3 = B + TMP0;
  float P4 = B - TMP0;
  float S1 = expf(-H * VAI);
  float S2 = expf(-EXT * VAI);
  float U1, U2, U3;
  if (IC == 0) {                                                // func:2393-2401
    U1 = B - C / ALBGRD(IB);
    U2 = B - C * ALBGRD(IB);
    U3 = F + C * ALBGRD(IB);
  } else {
    U1 = B - C / ALBGRI(IB);
    U2 = B - C * ALBGRI(IB);
    U3 = F + C * ALBGRI(IB);
  }
  TMP2 = U1 - AVMU * H;                                         // func:2402-2421
  float TMP3 = U1 + AVMU * H;
  float D1 = P1 * TMP2 / S1 - P2 * TMP3 * S1;
  float TMP4 = U2 + AVMU * H;
  float TMP5 = U2 - AVMU * H;
  float D2 = TMP4 / S1 - TMP5 * S1;
  float H1 = -D * P4 - C * F;
  float TMP6 = D - H1 * P3 / SIGMA;
  float TMP7 = (D - C - H1 / SIGMA * (U1 + TMP0)) * S2;
  float H2 = (TMP6 * TMP2 / S1 - P2 * TMP7) / D1;
  float H3 = -(TMP6 * TMP3 * S1 - P1 * TMP7) / D1;
  float H4 = -F * P3 - C * D;
  float TMP8 = H4 / SIGMA;
  float TMP9 = (U3 - TMP8 * (U2 - TMP0)) * S2;
  float H5 = -(TMP8 * TMP4 / S1 + TMP9) / D2;
  float H6 = (TMP8 * TMP5 * S1 + TMP9) / D2;
  float H7 = (C * TMP2) / (D1 * S1);
  float H8 = (-C * TMP3 * S1) / D1;
  float H9 = TMP4 / (D2 * S1);
  float H10 = (-TMP5 * S1) / D2;

  float FTDS, FTIS;
  if (IC == 0) {                                                // func:2426-2434
    FTDS = S2 * (1.0f - GAP) + GAP;
    FTIS = (H4 * S2 / SIGMA + H5 * S1 + H6 / S1) * (1.0f - GAP);
  } else {
    FTDS = 0.0f;
    FTIS = (H9 * S1 + H10 / S1) * (1.0f - KOPEN) + KOPEN;
  }
  FTD(IB) = FTDS;
  FTI(IB) = FTIS;

  float FRES, FREVEG, FREBAR;
  if (IC == 0) {                                                // func:2438-2450
    FRES = (H1 / SIGMA + H2 + H3) * (1.0f - GAP) + ALBGRD(IB) * GAP;
    FREVEG = (H1 / SIGMA + H2 + H3) * (1.0f - GAP);
    FREBAR = ALBGRD(IB) * GAP;
  } else {
    FRES = (H7 + H8) * (1.0f - KOPEN) + ALBGRI(IB) * KOPEN;
    FREVEG = (H7 + H8) * (1.0f - KOPEN) + ALBGRI(IB) * KOPEN;
    FREBAR = 0.0f;
  }
  FRE(IB) = FRES;
  FREV(IB) = FREVEG;
  FREG(IB) = FREBAR;

  FAB(IB) = 1.0f - FRE(IB) - (1.0f - ALBGRD(IB)) * FTD(IB) - (1.0f - ALBGRI(IB)) * FTI(IB);  // func:2454-2455
}

}  // namespace noahmp_oracle
