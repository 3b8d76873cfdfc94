// TEST INFRASTRUCTURE ONLY (see common.h).
// carbon, co2flux : core/module_noahmp_func.f90:6642-7025 (bvocflux, func:7028-7112, is never
// called by the reference -- the call at func:6749 is commented out -- and is not restated).
#include "physics.h"

namespace noahmp_oracle {

// func:6642-6751
void carbon(const Ctx& cx, int lutyp, float DT, const A4& ZSOIL, const A7& DZSNSO, const A7& STC,
            const A4& SMC, float TV, float TG, float PSN, float FOLN, float SMCMAX, float BTRAN,
            float APAR, float fveg, float IGS, float TROOT, int IST, float LAT, float& LFMASS,
            float& RTMASS, float& STMASS, float& WOOD, float& STBLCP, float& FASTCP, float& GPP,
            float& NPP, float& NEE, float& AUTORS, float& HETERS, float& TOTSC, float& TOTLB,
            float& XLAI, float& XSAI) {
  (void)TG; (void)APAR; (void)IST;
  const Tables& t = cx.tbl;
  if ((lutyp == t.ISWATER) || (lutyp == t.ISBARREN) || (lutyp == t.ISICE) || (lutyp == t.ISURBAN)) {
    XLAI = 0.0f; XSAI = 0.0f; GPP = 0.0f; NPP = 0.0f; NEE = 0.0f; AUTORS = 0.0f; HETERS = 0.0f;  // func:6709-6726
    TOTSC = 0.0f; TOTLB = 0.0f; LFMASS = 0.0f; RTMASS = 0.0f; STMASS = 0.0f; WOOD = 0.0f;
    STBLCP = 0.0f; FASTCP = 0.0f;
    return;
  }
  float LAPM = t.LK_SLA[lutyp] / 1000.0f;                       // func:6729
  float WSTRES = 1.0f - BTRAN;                                  // func:6733
  float WROOT = 0.0f;                                           // func:6735-6738
  const int NROOT = t.LK_NROOT[lutyp];
  for (int J = 1; J <= NROOT; ++J) WROOT = WROOT + SMC(J) / SMCMAX * DZSNSO(J) / (-ZSOIL(NROOT));

  co2flux(cx, lutyp, IGS, DT, DZSNSO, STC, PSN, TROOT, TV, WROOT, WSTRES, FOLN, LAPM, LAT, fveg, XLAI,
          XSAI, LFMASS, RTMASS, STMASS, FASTCP, STBLCP, WOOD, GPP, NPP, NEE, AUTORS, HETERS, TOTSC,
          TOTLB);                                               // func:6740-6747
}

// func:6754-7025
void co2flux(const Ctx& cx, int lutyp, float IGS, float DT, const A7& DZSNSO, const A7& STC,
             float PSN, float TROOT, float TV, float WROOT, float WSTRES, float FOLN, float LAPM,
             float LAT, float fveg, float& XLAI, float& XSAI, float& LFMASS, float& RTMASS,
             float& STMASS, float& FASTCP, float& STBLCP, float& WOOD, float& GPP, float& NPP,
             float& NEE, float& AUTORS, float& HETERS, float& TOTSC, float& TOTLB) {
  (void)DZSNSO; (void)TROOT; (void)LAT; (void)fveg;
  const Tables& t = cx.tbl;
  auto r = [](float x) { return expf(0.08f * (x - 298.16f)); };  // func:6882

  float RTOVRC = 2.0E-8f;                                       // func:6885-6891
  float RSWOODC = 3.0E-10f;
  float BF = 0.90f;
  float WSTRC = 100.0f;
  float LAIMIN = 0.05f;
  float XSAMIN = 0.01f;

  float SAPM = 3.0f * 0.001f;                                   // func:6893-6895
  float LFMSMN = LAIMIN / LAPM;
  float STMSMN = XSAMIN / SAPM;

  float RF;
  if (IGS == 0.0f) RF = 0.5f; else RF = 1.0f;                   // func:6899-6903

  float FNF = fmin2(FOLN / fmax2(1.0E-06f, t.LK_FOLNMX[lutyp]), 1.0f);  // func:6905-6912
  float TF = powf(t.LK_ARM[lutyp], (TV - 298.16f) / 10.0f);
  float RESP = t.LK_RMF25[lutyp] * TF * FNF * XLAI * RF * (1.0f - WSTRES);
  float RSLEAF = fmin2(LFMASS / DT, RESP * 12.0E-6f);
  float RSROOT = t.LK_RMR25[lutyp] * (RTMASS * 1.0E-3f) * TF * RF * 12.0E-6f;
  float RSSTEM = t.LK_RMS25[lutyp] * (STMASS * 1.0E-3f) * TF * RF * 12.0E-6f;
  float RSWOOD = RSWOODC * r(TV) * WOOD * t.LK_WDPOOL[lutyp];

  float CARBFX = PSN * 12.0E-6f;                                // func:6917

  float LEAFPT = expf(0.01f * (1.0f - expf(0.75f * XLAI)) * XLAI);  // func:6921-6926
  if (lutyp == t.ISEGBLF) LEAFPT = expf(0.01f * (1.0f - expf(0.50f * XLAI)) * XLAI);
  float NONLEF = 1.0f - LEAFPT;
  float STEMPT = XLAI / 10.0f;
  LEAFPT = LEAFPT - STEMPT;

  float WOODF;                                                  // func:6930-6937
  if (WOOD > 0.0f) WOODF = (1.0f - expf(-BF * (t.LK_WRRAT[lutyp] * RTMASS / WOOD)) / BF) * t.LK_WDPOOL[lutyp];
  else WOODF = 0.0f;
  float ROOTPT = NONLEF * (1.0f - WOODF);
  float WOODPT = NONLEF * WOODF;

  float LFTOVR = t.LK_LTOVRC[lutyp] * 1.0E-6f * LFMASS;         // func:6941-6944
  float STTOVR = t.LK_LTOVRC[lutyp] * 1.0E-6f * STMASS;
  float RTTOVR = RTOVRC * RTMASS;
  float WDTOVR = 9.5E-10f * WOOD;

  float SC = expf(-0.3f * fmax2(0.0f, TV - t.LK_TDLEF[lutyp])) * (LFMASS / 120.0f);  // func:6949-6952
  float SD = expf((WSTRES - 1.0f) * WSTRC);
  float DIELF = LFMASS * 1.0E-6f * (t.LK_DILEFW[lutyp] * SD + t.LK_DILEFC[lutyp] * SC);
  float DIEST = STMASS * 1.0E-6f * (t.LK_DILEFW[lutyp] * SD + t.LK_DILEFC[lutyp] * SC);

  float GRLEAF = fmax2(0.0f, t.LK_FRAGR[lutyp] * (LEAFPT * CARBFX - RSLEAF));  // func:6956-6959
  float GRSTEM = fmax2(0.0f, t.LK_FRAGR[lutyp] * (STEMPT * CARBFX - RSSTEM));
  float GRROOT = fmax2(0.0f, t.LK_FRAGR[lutyp] * (ROOTPT * CARBFX - RSROOT));
  float GRWOOD = fmax2(0.0f, t.LK_FRAGR[lutyp] * (WOODPT * CARBFX - RSWOOD));

  float ADDNPPLF = fmax2(0.0f, LEAFPT * CARBFX - GRLEAF - RSLEAF);  // func:6963-6966
  float ADDNPPST = fmax2(0.0f, STEMPT * CARBFX - GRSTEM - RSSTEM);
  if (TV < t.LK_TMIN[lutyp]) ADDNPPLF = 0.0f;
  if (TV < t.LK_TMIN[lutyp]) ADDNPPST = 0.0f;

  float LFDEL = (LFMASS - LFMSMN) / DT;                         // func:6971-6974
  float STDEL = (STMASS - STMSMN) / DT;
  DIELF = fmin2(DIELF, LFDEL + ADDNPPLF - LFTOVR);
  DIEST = fmin2(DIEST, STDEL + ADDNPPST - STTOVR);

  float NPPL = fmax2(ADDNPPLF, -LFDEL);                         // func:6978-6981
  float NPPS = fmax2(ADDNPPST, -STDEL);
  float NPPR = ROOTPT * CARBFX - RSROOT - GRROOT;
  float NPPW = WOODPT * CARBFX - RSWOOD - GRWOOD;

  LFMASS = LFMASS + (NPPL - LFTOVR - DIELF) * DT;               // func:6985-6993
  STMASS = STMASS + (NPPS - STTOVR - DIEST) * DT;
  RTMASS = RTMASS + (NPPR - RTTOVR) * DT;
  if (RTMASS < 0.0f) {
    RTTOVR = NPPR;
    RTMASS = 0.0f;
  }
  WOOD = (WOOD + (NPPW - WDTOVR) * DT) * t.LK_WDPOOL[lutyp];

  FASTCP = FASTCP + (RTTOVR + LFTOVR + STTOVR + WDTOVR + DIELF) * DT;  // func:6997

  float FST = powf(2.0f, (STC(1) - 283.16f) / 10.0f);           // func:6999-7001
  float FSW = WROOT / (0.20f + WROOT) * 0.23f / (0.23f + WROOT);
  float RSSOIL = FSW * FST * t.LK_MRP[lutyp] * fmax2(0.0f, FASTCP * 1.0E-3f) * 12.0E-6f;

  float STABLC = 0.1f * RSSOIL;                                 // func:7003-7005
  FASTCP = FASTCP - (RSSOIL + STABLC) * DT;
  STBLCP = STBLCP + STABLC * DT;

  float CFLUX = -CARBFX + RSLEAF + RSROOT + RSWOOD + RSSTEM + RSSOIL + GRLEAF + GRROOT + GRWOOD;  // func:7008
  (void)CFLUX;

  GPP = CARBFX;                                                 // func:7012-7019
  NPP = NPPL + NPPW + NPPR;
  AUTORS = RSROOT + RSWOOD + RSLEAF + GRLEAF + GRROOT + GRWOOD;
  HETERS = RSSOIL;
  NEE = (AUTORS + HETERS - GPP) * 44.0f / 12.0f;
  TOTSC = FASTCP + STBLCP;
  TOTLB = LFMASS + RTMASS + WOOD;

  XLAI = fmax2(LFMASS * LAPM, LAIMIN);                          // func:7022-7023
  XSAI = fmax2(STMASS * SAPM, XSAMIN);
}

}  // namespace noahmp_oracle
