// TEST INFRASTRUCTURE ONLY (see common.h).
// energy + thermoprop + csnow + tdfcnd : core/module_noahmp_func.f90:735-1595
#include "physics.h"

namespace noahmp_oracle {

// func:735-1338
void energy(const Ctx& cx, ColumnIO& c, float RHOAIR, float QAIR, float THAIR, float EAIR,
            const A2& SOLAD, const A2& SOLAI, float IGS, float HTOP, float elai, float esai,
            A7& DZSNSO, EnergyOut& o) {
  const Tables& t = cx.tbl;
  const Options& op = cx.opt;
  const int sltyp = c.sltyp, lutyp = c.lutyp, IST = c.IST, ICE = c.ICE, ISNOW = c.ISNOW;
  const float DT = c.DT, SFCPRS = c.SFCPRS, SFCTMP = c.SFCTMP, LWDN = c.LWDN, UU = c.UU, VV = c.VV;
  const float ZREF = c.ZLVL, fveg = c.fveg;
  const float PSIWLT = -150.0f;                                 // func:1017
  const float Z0 = 0.01f;                                       // func:1018

  float TAUXV = 0.0f, TAUYV = 0.0f;                             // func:1021-1038
  c.IRC = 0.0f; c.SHC = 0.0f; c.IRG = 0.0f; c.SHG = 0.0f; c.EVG = 0.0f; c.EVC = 0.0f; c.TR = 0.0f;
  c.GHV = 0.0f;
  float PSNSUN = 0.0f, PSNSHA = 0.0f;
  c.T2MV = 0.0f; c.Q2V = 0.0f; c.CHV = 0.0f; c.CHLEAF = 0.0f; c.CHUC = 0.0f; c.CHV2 = 0.0f;

  float UR = fmax2(sqrtf(UU * UU + VV * VV), 1.0f);             // func:1041

  float VAI = elai + esai;                                      // func:1044-1046
  bool VEG = false;
  if (VAI > 0.0f) VEG = true;

  c.FSNO = 0.0f;                                                // func:1049-1054
  if (c.snowh > 0.0f) {
    float BDSNO = c.sneqv / c.snowh;
    float FMELT = powf(BDSNO / 100.0f, t.KK_MLTFCT);
    c.FSNO = tanhf(c.snowh / (2.5f * Z0 * FMELT));
  }

  float Z0MG;                                                   // func:1057-1065
  if (IST == 2) {
    if (c.TG <= TFRZ) Z0MG = 0.01f * (1.0f - c.FSNO) + c.FSNO * t.KK_Z0SNO;
    else Z0MG = 0.01f;
  } else {
    Z0MG = Z0 * (1.0f - c.FSNO) + c.FSNO * t.KK_Z0SNO;
  }

  float ZPDG = c.snowh;                                         // func:1069-1077
  float Z0M, ZPD;
  if (VEG) {
    Z0M = t.LK_Z0MVT[lutyp];
    ZPD = 0.65f * HTOP;
    if (c.snowh > ZPD) ZPD = c.snowh;
  } else {
    Z0M = Z0MG;
    ZPD = ZPDG;
  }

  float ZLVL = fmax2(ZPD, HTOP) + ZREF;                         // func:1079-1080
  if (ZPDG >= ZLVL) ZLVL = ZPDG + ZREF;

  float CWP = t.LK_CWPVT[lutyp];                                // func:1084

  A7 DF, HCPCT, FACT;
  for (int i = 0; i < 7; ++i) { DF.v[i] = 0.0f; HCPCT.v[i] = 0.0f; FACT.v[i] = 0.0f; }
  thermoprop(cx, sltyp, ISNOW, IST, DZSNSO, DT, c.snowh, c.SNICE, c.SNLIQ, t.KK_CSOIL, c.SMC,
             c.soilwat, c.STC, lutyp, DF, HCPCT, o.SNICEV, o.SNLIQV, o.epore, FACT);  // func:1087-1092

  float FSUN, LAISUN, LAISHA, PARSUN, PARSHA;
  radiation(cx, lutyp, IST, c.ISC, ICE, c.SNEQVO, c.sneqv, DT, c.COSZ, c.snowh, c.TG, c.TV, c.FSNO,
            c.QSNOW, c.FWET, elai, esai, c.SMC, SOLAD, SOLAI, fveg, c.ALBOLD, c.TAUSS, FSUN, LAISUN,
            LAISHA, PARSUN, PARSHA, c.SAV, c.SAG, c.FSR, c.FSA, o.FSRV, o.FSRG, c.BGAP,
            c.WGAP);                                            // func:1095-1103

  float EMV = 1.0f - expf(-(elai + esai) / 1.0f);               // func:1106
  float EMG;
  if (ICE == 1) EMG = 0.98f * (1.0f - c.FSNO) + 1.0f * c.FSNO;  // func:1107-1113
  else if (IST == 1) EMG = t.KK_EMSSOIL * (1.0f - c.FSNO) + 1.0f * c.FSNO;
  else EMG = t.KK_EMSLAKE * (1.0f - c.FSNO) + 1.0f * c.FSNO;

  o.BTRAN = 0.0f;                                               // func:1117
  for (int IZ = 1; IZ <= NSOIL; ++IZ) o.BTRANI(IZ) = 0.0f;      // (undefined in Fortran beyond NROOT)
  const int NROOT = t.LK_NROOT[lutyp];
  if (IST == 1) {                                               // func:1119-1140
    for (int IZ = 1; IZ <= NROOT; ++IZ) {
      float GX = 0.0f, PSI;
      if (op.opt_btr == 1) {
        GX = (c.soilwat(IZ) - t.LK_SMCWLT[sltyp]) / (t.LK_SMCREF[sltyp] - (t.LK_SMCWLT[sltyp]));
      }
      if (op.opt_btr == 2) {
        PSI = fmax2(PSIWLT, -t.LK_PSISAT[sltyp] *
                                powf(fmax2(0.01f, c.soilwat(IZ)) / t.LK_SMCMAX[sltyp], -t.LK_BEXP[sltyp]));
        GX = (1.0f - PSI / PSIWLT) / (1.0f + t.LK_PSISAT[sltyp] / PSIWLT);
      }
      if (op.opt_btr == 3) {
        PSI = fmax2(PSIWLT, -t.LK_PSISAT[sltyp] *
                                powf(fmax2(0.01f, c.soilwat(IZ)) / t.LK_SMCMAX[sltyp], -t.LK_BEXP[sltyp]));
        GX = 1.0f - expf(-5.8f * (logf(PSIWLT / PSI)));
      }
      GX = fmin2(1.0f, fmax2(0.0f, GX));                        // func:1133
      o.BTRANI(IZ) = fmax2(MPE, DZSNSO(IZ) / (-c.ZSOIL(NROOT)) * GX);  // func:1134
      o.BTRAN = o.BTRAN + o.BTRANI(IZ);
    }
    o.BTRAN = fmax2(MPE, o.BTRAN);                              // func:1137
    for (int IZ = 1; IZ <= NROOT; ++IZ) o.BTRANI(IZ) = o.BTRANI(IZ) / o.BTRAN;  // func:1139
  }

  float BEVAP = fmax2(0.0f, c.soilwat(1) / t.LK_SMCMAX[sltyp]);  // func:1143
  (void)BEVAP;
  float RSURF, RHSUR;
  if (IST == 2) {                                               // func:1144-1164
    RSURF = 1.0f;
    RHSUR = 1.0f;
  } else {
    float L_RSURF = (-c.ZSOIL(1)) *
                    (expf(pow5i(1.0f - fmin2(1.0f, c.soilwat(1) / t.LK_SMCMAX[sltyp]))) - 1.0f) /
                    (2.71828f - 1.0f);                          // func:1152
    float D_RSURF = 2.2E-5f * t.LK_SMCMAX[sltyp] * t.LK_SMCMAX[sltyp] *
                    powf(1.0f - t.LK_SMCWLT[sltyp] / t.LK_SMCMAX[sltyp],
                         2.0f + 3.0f / t.LK_BEXP[sltyp]);       // func:1153-1154
    RSURF = L_RSURF / D_RSURF;                                  // func:1155
    if (c.soilwat(1) < 0.01f && c.snowh == 0.0f) RSURF = 1.0E6f;  // func:1161
    float PSI = -t.LK_PSISAT[sltyp] *
                powf(fmax2(0.01f, c.soilwat(1)) / t.LK_SMCMAX[sltyp], -t.LK_BEXP[sltyp]);  // func:1162
    RHSUR = c.FSNO + (1.0f - c.FSNO) * expf(PSI * GRAV / (RVAP * c.TG));                   // func:1163
  }

  if (lutyp == t.ISURBAN && c.snowh == 0.0f) RSURF = 1.0E6f;    // func:1167-1169

  if (c.TV > TFRZ) { o.LATHEAV = HVAP; o.frozen_canopy = false; }  // func:1173-1179
  else { o.LATHEAV = HSUB; o.frozen_canopy = true; }
  float GAMMAV = CPAIR * SFCPRS / (0.622f * o.LATHEAV);         // func:1180

  if (c.TG > TFRZ) { o.LATHEAG = HVAP; o.frozen_ground = false; }  // func:1182-1188
  else { o.LATHEAG = HSUB; o.frozen_ground = true; }
  float GAMMAG = CPAIR * SFCPRS / (0.622f * o.LATHEAG);         // func:1189

  float CMV = 0.0f;
  if (VEG && fveg > 0.0f) {                                     // func:1200-1224
    c.TGV = c.TG;
    CMV = c.CM;
    c.CHV = c.CH;
    vege_flux(cx, c, ISNOW, lutyp, DT, c.SAV, c.SAG, LWDN, UR, UU, VV, SFCTMP, THAIR, QAIR, EAIR,
              RHOAIR, c.snowh, VAI, GAMMAV, GAMMAG, c.FWET, LAISUN, LAISHA, CWP, DZSNSO, HTOP, ZLVL,
              ZPD, Z0M, fveg, Z0MG, EMV, EMG, c.CANLIQ, c.CANICE, c.STC, DF, c.RSSUN, c.RSSHA, RSURF,
              o.LATHEAV, o.LATHEAG, PARSUN, PARSHA, IGS, c.FOLN, c.CO2AIR, c.O2AIR, o.BTRAN, SFCPRS,
              RHSUR, c.EAH, c.TAH, c.TV, c.TGV, CMV, c.CHV, TAUXV, TAUYV, c.IRG, c.IRC, c.SHG, c.SHC,
              c.EVG, c.EVC, c.TR, c.GHV, c.T2MV, PSNSUN, PSNSHA, c.QSFC, c.PSFC, c.Q2V, c.CHV2,
              c.CHLEAF, c.CHUC);
  }

  c.TGB = c.TG;                                                 // func:1226-1239
  float CMB = c.CM;
  c.CHB = c.CH;
  float TAUXB, TAUYB;
  bare_flux(cx, c, lutyp, DT, ISNOW, c.SAG, LWDN, UR, UU, VV, SFCTMP, THAIR, QAIR, EAIR, RHOAIR,
            c.snowh, DZSNSO, ZLVL, ZPDG, Z0MG, EMG, c.STC, DF, RSURF, o.LATHEAG, GAMMAG, RHSUR, c.TGB,
            CMB, c.CHB, TAUXB, TAUYB, c.IRB, c.SHB, c.EVB, c.GHB, c.T2MB, c.QSFC, c.PSFC, SFCPRS,
            c.Q2B, c.CHB2);

  if (VEG && fveg > 0.0f) {                                     // func:1246-1282
    o.TAUX = fveg * TAUXV + (1.0f - fveg) * TAUXB;
    o.TAUY = fveg * TAUYV + (1.0f - fveg) * TAUYB;
    c.FIRA = fveg * c.IRG + (1.0f - fveg) * c.IRB + c.IRC;
    c.FSH = fveg * c.SHG + (1.0f - fveg) * c.SHB + c.SHC;
    c.FGEV = fveg * c.EVG + (1.0f - fveg) * c.EVB;
    c.SSOIL = fveg * c.GHV + (1.0f - fveg) * c.GHB;
    c.FCEV = c.EVC;
    c.FCTR = c.TR;
    c.TG = fveg * c.TGV + (1.0f - fveg) * c.TGB;
    o.T2M = fveg * c.T2MV + (1.0f - fveg) * c.T2MB;
    o.TS = fveg * c.TV + (1.0f - fveg) * c.TGB;
    c.CM = fveg * CMV + (1.0f - fveg) * CMB;
    c.CH = fveg * c.CHV + (1.0f - fveg) * c.CHB;
    o.Q1 = fveg * (c.EAH * 0.622f / (SFCPRS - 0.378f * c.EAH)) + (1.0f - fveg) * c.QSFC;
    o.Q2E = fveg * c.Q2V + (1.0f - fveg) * c.Q2B;
  } else {
    o.TAUX = TAUXB;
    o.TAUY = TAUYB;
    c.FIRA = c.IRB;
    c.FSH = c.SHB;
    c.FGEV = c.EVB;
    c.SSOIL = c.GHB;
    c.TG = c.TGB;
    o.T2M = c.T2MB;
    c.FCEV = 0.0f;
    c.FCTR = 0.0f;
    o.TS = c.TG;
    c.CM = CMB;
    c.CH = c.CHB;
    o.Q1 = c.QSFC;
    o.Q2E = c.Q2B;
    c.RSSUN = 0.0f;
    c.RSSHA = 0.0f;
    c.TGV = c.TGB;
    c.CHV = c.CHB;
  }

  float FIRE = LWDN + c.FIRA;                                   // func:1284
  if (FIRE <= 0.0f) c.flags |= FLAG_FIRE;                       // func:1286-1292

  c.EMISSI = fveg * (EMG * (1.0f - EMV) + EMV + EMV * (1.0f - EMV) * (1.0f - EMG)) +
             (1.0f - fveg) * EMG;                               // func:1295-1296

  c.TRAD = powf((FIRE - (1.0f - c.EMISSI) * LWDN) / (c.EMISSI * SB), 0.25f);  // func:1302

  c.APAR = PARSUN * LAISUN + PARSHA * LAISHA;                   // func:1307
  c.PSN = PSNSUN * LAISUN + PSNSHA * LAISHA;                    // func:1308

  tsnosoi(cx, ISNOW, c.TBOT, c.ZSNSO, c.SSOIL, DF, HCPCT, t.KK_ZBOT, DT, c.snowh, c.STC);  // func:1311-1315

  if (op.opt_stc == 2) {                                        // func:1318-1330
    if (c.snowh > 0.05f && c.TG > TFRZ) {
      c.TGV = TFRZ;
      c.TGB = TFRZ;
      if (VEG && fveg > 0.0f) {
        c.TG = fveg * c.TGV + (1.0f - fveg) * c.TGB;
        o.TS = fveg * c.TV + (1.0f - fveg) * c.TGB;
      } else {
        c.TG = c.TGB;
        o.TS = c.TGB;
      }
    }
  }

  phasechange(cx, sltyp, ISNOW, DT, FACT, DZSNSO, HCPCT, IST, c.STC, c.SNICE, c.SNLIQ, c.sneqv,
              c.snowh, c.SMC, c.soilwat, o.QMELT, o.IMELT, c.PONDING);  // func:1333-1337
}

// func:1341-1445
void thermoprop(const Ctx& cx, int sltyp, int ISNOW, int IST, const A7& DZSNSO, float DT, float snowh,
                const A3& SNICE, const A3& SNLIQ, float CSOIL, const A4& SMC, const A4& soilwat,
                const A7& STC, int lutyp, A7& DF, A7& HCPCT, A3& SNICEV, A3& SNLIQV, A3& epore,
                A7& FACT) {
  const Tables& t = cx.tbl;
  A3 CVSNO, TKSNO;
  for (int i = 0; i < 3; ++i) { CVSNO.v[i] = 0.0f; TKSNO.v[i] = 0.0f; }
  A4 soilice;
  csnow(ISNOW, SNICE, SNLIQ, DZSNSO, TKSNO, CVSNO, SNICEV, SNLIQV, epore);  // func:1388-1389

  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {                     // func:1391-1394
    DF(IZ) = TKSNO(IZ);
    HCPCT(IZ) = CVSNO(IZ);
  }

  for (int IZ = 1; IZ <= NSOIL; ++IZ) {                         // func:1398-1403
    soilice(IZ) = SMC(IZ) - soilwat(IZ);
    HCPCT(IZ) = soilwat(IZ) * CWAT + (1.0f - t.LK_SMCMAX[sltyp]) * CSOIL +
                (t.LK_SMCMAX[sltyp] - SMC(IZ)) * CPAIR + soilice(IZ) * CICE;
    tdfcnd(cx, sltyp, DF(IZ), SMC(IZ), soilwat(IZ));
  }

  if (lutyp == t.ISURBAN) {                                     // func:1405-1409
    for (int IZ = 1; IZ <= NSOIL; ++IZ) DF(IZ) = 3.24f;
  }

  if (IST == 2) {                                               // func:1420-1430
    for (int IZ = 1; IZ <= NSOIL; ++IZ) {
      if (STC(IZ) > TFRZ) {
        HCPCT(IZ) = CWAT;
        DF(IZ) = TKWAT;
      } else {
        HCPCT(IZ) = CICE;
        DF(IZ) = TKICE;
      }
    }
  }

  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) FACT(IZ) = DT / (HCPCT(IZ) * DZSNSO(IZ));  // func:1434-1436

  if (ISNOW == 0) {                                             // func:1440-1444
    DF(1) = (DF(1) * DZSNSO(1) + 0.35f * snowh) / (snowh + DZSNSO(1));
  } else {
    DF(1) = (DF(1) * DZSNSO(1) + DF(0) * DZSNSO(0)) / (DZSNSO(0) + DZSNSO(1));
  }
}

// func:1448-1497
void csnow(int ISNOW, const A3& SNICE, const A3& SNLIQ, const A7& DZSNSO, A3& TKSNO, A3& CVSNO,
           A3& SNICEV, A3& SNLIQV, A3& epore) {
  A3 BDSNOI;
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {                     // func:1475-1479
    SNICEV(IZ) = fmin2(1.0f, SNICE(IZ) / (DZSNSO(IZ) * DENICE));
    epore(IZ) = 1.0f - SNICEV(IZ);
    SNLIQV(IZ) = fmin2(epore(IZ), SNLIQ(IZ) / (DZSNSO(IZ) * DENWAT));
  }
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {                     // func:1481-1485
    BDSNOI(IZ) = (SNICE(IZ) + SNLIQ(IZ)) / DZSNSO(IZ);
    CVSNO(IZ) = CICE * SNICEV(IZ) + CWAT * SNLIQV(IZ);
  }
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {                     // func:1489-1495
    TKSNO(IZ) = 3.2217E-6f * powf(BDSNOI(IZ), 2.0f);
  }
}

// func:1500-1595
void tdfcnd(const Ctx& cx, int sltyp, float& DF, float SMC, float soilwat) {
  const Tables& t = cx.tbl;
  float SATRATIO = SMC / t.LK_SMCMAX[sltyp];                    // func:1552
  float THKW = 0.57f;
  float THKO = 2.0f;
  float THKQTZ = 7.7f;
  float THKS = (powf(THKQTZ, t.LK_QUARTZ[sltyp])) * (powf(THKO, 1.0f - t.LK_QUARTZ[sltyp]));  // func:1561
  float XUNFROZ = soilwat / SMC;                                // func:1564
  float XU = XUNFROZ * t.LK_SMCMAX[sltyp];                      // func:1566
  float THKSAT = powf(THKS, 1.0f - t.LK_SMCMAX[sltyp]) * powf(TKICE, t.LK_SMCMAX[sltyp] - XU) *
                 powf(THKW, XU);                                // func:1569
  float GAMMD = (1.0f - t.LK_SMCMAX[sltyp]) * 2700.0f;          // func:1572
  float THKDRY = (0.135f * GAMMD + 64.7f) / (2700.0f - 0.947f * GAMMD);  // func:1574
  float AKE;
  if ((soilwat + 0.0005f) < SMC) {                              // func:1576
    AKE = SATRATIO;                                             // frozen
  } else {
    if (SATRATIO > 0.1f) AKE = log10f(SATRATIO) + 1.0f;         // func:1584-1589
    else AKE = 0.0f;
  }
  DF = AKE * (THKSAT - THKDRY) + THKDRY;                        // func:1593
}

}  // namespace noahmp_oracle
