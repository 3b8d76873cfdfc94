// TEST INFRASTRUCTURE ONLY (see common.h).
// noahmp_sflx + atm + phenology + error : core/module_noahmp_func.f90:66-732
#include "physics.h"

namespace noahmp_oracle {

// func:479-531
void atm(float SFCPRS, float SFCTMP, float Q2, float PRCP, float SOLDN, float COSZ, float& THAIR,
         float& QAIR, float& EAIR, float& RHOAIR, float& QPRECC, float& QPRECL, A2& SOLAD, A2& SOLAI,
         float& SWDOWN) {
  float PAIR = SFCPRS;                                          // func:508
  THAIR = SFCTMP * powf(SFCPRS / PAIR, RAIR / CPAIR);           // func:509
  QAIR = Q2;                                                    // func:512
  EAIR = QAIR * SFCPRS / (0.622f + 0.378f * QAIR);              // func:514
  RHOAIR = (SFCPRS - 0.378f * EAIR) / (RAIR * SFCTMP);          // func:515
  QPRECC = 0.10f * PRCP;                                        // func:517
  QPRECL = 0.90f * PRCP;                                        // func:518
  if (COSZ <= 0.0f) SWDOWN = 0.0f; else SWDOWN = SOLDN;         // func:520-524
  SOLAD(1) = SWDOWN * 0.7f * 0.5f;                              // func:526-529
  SOLAD(2) = SWDOWN * 0.7f * 0.5f;
  SOLAI(1) = SWDOWN * 0.3f * 0.5f;
  SOLAI(2) = SWDOWN * 0.3f * 0.5f;
}

// func:534-630
void phenology(const Ctx& cx, int lutyp, float snowh, float TV, float LAT, int YEARLEN, float JULIAN,
               float& LAI, float& SAI, float TROOT, float& HTOP, float& elai, float& esai, float& IGS) {
  (void)TROOT;
  const Tables& t = cx.tbl;
  const int opt_veg = cx.opt.opt_veg;
  if (opt_veg == 1 || opt_veg == 3 || opt_veg == 4) {           // func:578
    float DAY;
    if (LAT >= 0.0f) {
      DAY = JULIAN;                                             // func:582
    } else {
      DAY = fmodf(JULIAN + (0.5f * (float)YEARLEN), (float)YEARLEN);  // func:585
    }
    float T = 12.0f * DAY / (float)YEARLEN;                     // func:588
    int IT1 = (int)(T + 0.5f);                                  // func:589 (real->int truncation)
    int IT2 = IT1 + 1;
    float WT1 = ((float)IT1 + 0.5f) - T;                        // func:591
    float WT2 = 1.0f - WT1;
    if (IT1 < 1) IT1 = 12;
    if (IT2 > 12) IT2 = 1;
    LAI = WT1 * t.LK_LAI12M[IT1][lutyp] + WT2 * t.LK_LAI12M[IT2][lutyp];  // func:596
    SAI = WT1 * t.LK_SAI12M[IT1][lutyp] + WT2 * t.LK_SAI12M[IT2][lutyp];  // func:597
  }
  if (SAI < 0.05f) SAI = 0.0f;                                  // func:599
  if (LAI < 0.05f || SAI == 0.0f) LAI = 0.0f;                   // func:600
  if (lutyp == t.ISWATER || lutyp == t.ISBARREN || lutyp == t.ISICE || lutyp == t.ISURBAN) {
    LAI = 0.0f;                                                 // func:602-605
    SAI = 0.0f;
  }
  float DB = fmin2(fmax2(snowh - t.LK_HVB[lutyp], 0.0f), t.LK_HVT[lutyp] - t.LK_HVB[lutyp]);  // func:609
  float FB = DB / fmax2(1.0E-06f, t.LK_HVT[lutyp] - t.LK_HVB[lutyp]);                          // func:610
  if (t.LK_HVT[lutyp] > 0.0f && t.LK_HVT[lutyp] <= 1.0f) {      // func:612
    float snowhc = t.LK_HVT[lutyp] * expf(-snowh / 0.2f);       // func:613
    FB = fmin2(snowh, snowhc) / snowhc;                         // func:614
  }
  elai = LAI * (1.0f - FB);                                     // func:617
  esai = SAI * (1.0f - FB);
  if (esai < 0.05f) esai = 0.0f;                                // func:619
  if (elai < 0.05f || esai == 0.0f) elai = 0.0f;                // func:620
  if (TV > t.LK_TMIN[lutyp]) IGS = 1.0f; else IGS = 0.0f;       // func:622-626
  HTOP = t.LK_HVT[lutyp];                                       // func:628
}

// func:633-732 -- the two fatal closures become flag bits, ERRWAT is returned
void error(ColumnIO& c, float SWDOWN, float BEG_WB, const A7& DZSNSO) {
  float ERRSW = SWDOWN - (c.FSA + c.FSR);                       // func:688
  c.ERRSW = ERRSW;
  if (fabsf(ERRSW) > 0.01f) c.flags |= FLAG_ERRSW;              // func:691-710
  float ERRENG = c.SAV + c.SAG - (c.FIRA + c.FSH + c.FCEV + c.FGEV + c.FCTR + c.SSOIL);  // func:712
  c.ERRENG = ERRENG;
  if (fabsf(ERRENG) > 0.01f) c.flags |= FLAG_ERRENG;            // func:715-721
  if (c.IST == 1) {                                             // func:723
    float END_WB = c.CANLIQ + c.CANICE + c.sneqv + c.WA;        // func:724
    for (int IZ = 1; IZ <= NSOIL; ++IZ) END_WB = END_WB + c.SMC(IZ) * DZSNSO(IZ) * 1000.0f;
    c.ERRWAT = END_WB - BEG_WB - (c.PRCP - c.ECAN - c.ETRAN - c.EDIR - c.runsrf - c.runsub) * c.DT;
  } else {
    c.ERRWAT = 0.0f;                                            // func:730
  }
}

// func:66-476
void noahmp_sflx(const Ctx& cx, ColumnIO& c) {
  const Tables& t = cx.tbl;
  const Options& o = cx.opt;
  c.flags = 0;
  c.iter_vege = 0;
  c.iter_stom_sun = 0;
  c.iter_stom_sha = 0;
  c.niter_soil = 0;

  c.NEE = 0.0f;                                                 // func:310-312
  c.NPP = 0.0f;
  c.GPP = 0.0f;

  float THAIR, QAIR, EAIR, RHOAIR, QPRECC, QPRECL, SWDOWN;
  A2 SOLAD, SOLAI;
  atm(c.SFCPRS, c.SFCTMP, c.Q2, c.PRCP, c.SOLDN, c.COSZ, THAIR, QAIR, EAIR, RHOAIR, QPRECC, QPRECL,
      SOLAD, SOLAI, SWDOWN);                                    // func:316-318

  A7 DZSNSO;
  for (int i = 0; i < 7; ++i) DZSNSO.v[i] = 0.0f;               // undefined above ISNOW+1 in Fortran
  for (int IZ = c.ISNOW + 1; IZ <= NSOIL; ++IZ) {               // func:322-328
    if (IZ == c.ISNOW + 1) DZSNSO(IZ) = -c.ZSNSO(IZ);
    else DZSNSO(IZ) = c.ZSNSO(IZ - 1) - c.ZSNSO(IZ);
  }

  float TROOT = 0.0f;                                           // func:332-335
  const int NROOT = t.LK_NROOT[c.lutyp];
  for (int IZ = 1; IZ <= NROOT; ++IZ) TROOT = TROOT + c.STC(IZ) * DZSNSO(IZ) / (-c.ZSOIL(NROOT));

  float BEG_WB = 0.0f;
  if (c.IST == 1) {                                             // func:339-344
    BEG_WB = c.CANLIQ + c.CANICE + c.sneqv + c.WA;
    for (int IZ = 1; IZ <= NSOIL; ++IZ) BEG_WB = BEG_WB + c.SMC(IZ) * DZSNSO(IZ) * 1000.0f;
  }

  float HTOP, elai, esai, IGS;
  phenology(cx, c.lutyp, c.snowh, c.TV, c.LAT, c.YEARLEN, c.JULIAN, c.LAI, c.SAI, TROOT, HTOP, elai,
            esai, IGS);                                         // func:348-350

  float fveg = 0.0f;
  if (o.opt_veg == 1) {                                         // func:366-378
    fveg = c.SHDFAC;
    if (fveg <= 0.01f) fveg = 0.01f;
  } else if (o.opt_veg == 2 || o.opt_veg == 3) {
    fveg = 1.0f - expf(-0.52f * (c.LAI + c.SAI));
    if (fveg <= 0.01f) fveg = 0.01f;
  } else if (o.opt_veg == 4 || o.opt_veg == 5) {
    fveg = c.SHDMAX;
    if (fveg <= 0.01f) fveg = 0.01f;
  } else {
    c.flags |= FLAG_OPTVEG;
  }
  if (c.lutyp == t.ISURBAN || c.lutyp == t.ISBARREN) fveg = 0.0f;  // func:379
  if (elai + esai == 0.0f) fveg = 0.0f;                            // func:380
  c.fveg = fveg;

  EnergyOut eo;
  std::memset(&eo, 0, sizeof(eo));
  energy(cx, c, RHOAIR, QAIR, THAIR, EAIR, SOLAD, SOLAI, IGS, HTOP, elai, esai, DZSNSO, eo);  // func:388-411

  A4 soilice;
  for (int IZ = 1; IZ <= NSOIL; ++IZ) soilice(IZ) = fmax2(0.0f, c.SMC(IZ) - c.soilwat(IZ));   // func:414
  c.SNEQVO = c.sneqv;                                           // func:415

  float QVAP = fmax2(c.FGEV / eo.LATHEAG, 0.0f);                // func:417
  float QDEW = fabsf(fmin2(c.FGEV / eo.LATHEAG, 0.0f));         // func:418
  c.EDIR = QVAP - QDEW;                                         // func:419

  float CMC, QIN, QDIS;
  water(cx, c, eo.IMELT, QPRECC, QPRECL, elai, esai, QVAP, QDEW, eo.BTRANI, c.PONDING,
        eo.frozen_canopy, eo.frozen_ground, soilice, DZSNSO, CMC, QIN, QDIS);  // func:423-433

  if (o.opt_veg == 2 || o.opt_veg == 5) {                       // func:439-447
    float AUTORS, HETERS, TOTSC, TOTLB;
    carbon(cx, c.lutyp, c.DT, c.ZSOIL, DZSNSO, c.STC, c.SMC, c.TV, c.TG, c.PSN, c.FOLN,
           t.LK_SMCMAX[c.sltyp], eo.BTRAN, c.APAR, c.fveg, IGS, TROOT, c.IST, c.LAT, c.LFMASS, c.RTMASS,
           c.STMASS, c.WOOD, c.STBLCP, c.FASTCP, c.GPP, c.NPP, c.NEE, AUTORS, HETERS, TOTSC, TOTLB,
           c.LAI, c.SAI);
  }

  error(c, SWDOWN, BEG_WB, DZSNSO);                             // func:451-456

  float QFX = c.ETRAN + c.ECAN + c.EDIR;                        // func:459
  if (c.lutyp == t.ISURBAN) {                                   // func:460-463
    c.QSFC = (QFX / RHOAIR * c.CH) + QAIR;
    c.Q2B = c.QSFC;
  }

  if (c.snowh <= 1.0E-6f || c.sneqv <= 1.0E-3f) {               // func:465-468
    c.snowh = 0.0f;
    c.sneqv = 0.0f;
  }

  if (SWDOWN != 0.0f) c.ALBEDO = c.FSR / SWDOWN;                // func:470-474
  else c.ALBEDO = -999.9f;

  // added instrumentation: non-finite prognostics
  bool bad = !std::isfinite(c.TV) || !std::isfinite(c.TG) || !std::isfinite(c.sneqv) ||
             !std::isfinite(c.ZWT) || !std::isfinite(c.WA);
  for (int IZ = 1; IZ <= NSOIL; ++IZ) bad = bad || !std::isfinite(c.STC(IZ)) || !std::isfinite(c.SMC(IZ));
  if (bad) c.flags |= FLAG_NAN;
}

}  // namespace noahmp_oracle
