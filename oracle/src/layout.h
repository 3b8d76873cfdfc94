// TEST INFRASTRUCTURE ONLY (see common.h).
// Structure-of-arrays plane indices of the batch interface.  This is the oracle's OWN statement of
// the layout; tests/test_abi_layout.py checks it against the product library's exported names so
// the two can never drift silently.
#pragma once

namespace noahmp_oracle {

enum StatePlane {
  S_ALBOLD = 0, S_SNEQVO, S_EAH, S_FWET, S_CANLIQ, S_CANICE, S_TV, S_TG, S_QSNOW, S_ISNOW /*int32 bits*/,
  S_SNOWH, S_SNEQV, S_ZWT, S_WA, S_WT, S_TAUSS,
  S_STC = 16,      // 7 planes, Fortran index -2..4
  S_SOILWAT = 23,  // 4 planes (liquid soil moisture, the reference's SH2O)
  S_SMC = 27,      // 4
  S_ZSNSO = 31,    // 7
  S_SNICE = 38,    // 3, Fortran index -2..0
  S_SNLIQ = 41,    // 3
  NSTATE_CORE = 44,
  // extended state: only live for non-default options (SURVEY App. A group "s")
  S_QSFC = 44, S_TAH, S_CM, S_CH, S_WSLAKE, S_LFMASS, S_RTMASS, S_STMASS, S_WOOD, S_STBLCP, S_FASTCP,
  S_LAI, S_SAI,
  NSTATE = 57
};

enum ForcingPlane { F_SFCTMP = 0, F_SFCPRS, F_UU, F_VV, F_Q2, F_SOLDN, F_LWDN, F_PRCP, F_COSZ, NFORCING = 9 };

enum StaticPlane { C_LAT = 0, C_SHDFAC, C_TBOT, C_SHDMAX, NSTATIC_F = 4 };

// packed int32 per column: lutyp | sltyp<<8 | slptyp<<16 | ISC<<24 | (IST-1)<<28 | ICE<<29
static inline int pack_types(int lutyp, int sltyp, int slptyp, int isc, int ist, int ice) {
  return (lutyp & 0xff) | ((sltyp & 0xff) << 8) | ((slptyp & 0xff) << 16) | ((isc & 0xf) << 24) |
         (((ist - 1) & 1) << 28) | ((ice & 1) << 29);
}

// diagnostics planes: the 58 outputs of noahmp_sflx in the order of func:82-91, then budget residuals
enum DiagPlane {
  D_FSA = 0, D_FSR, D_FIRA, D_FSH, D_SSOIL, D_FCEV, D_FGEV, D_FCTR, D_ECAN, D_ETRAN, D_EDIR, D_TRAD,
  D_TGB, D_TGV, D_T2MV, D_T2MB, D_Q2V, D_Q2B, D_RUNSRF, D_RUNSUB, D_APAR, D_PSN, D_SAV, D_SAG, D_FSNO,
  D_NEE, D_GPP, D_NPP, D_FVEG, D_ALBEDO, D_QSNBOT, D_PONDING, D_PONDING1, D_PONDING2, D_RSSUN, D_RSSHA,
  D_BGAP, D_WGAP, D_CHV, D_CHB, D_EMISSI, D_SHG, D_SHC, D_SHB, D_EVG, D_EVB, D_GHV, D_GHB, D_IRG,
  D_IRC, D_IRB, D_TR, D_EVC, D_CHLEAF, D_CHUC, D_CHV2, D_CHB2, D_FPICE,
  D_ERRWAT = 58, D_ERRSW, D_ERRENG,
  D_ITER = 61,   // int32 bits: iter_vege | stom_sun<<8 | stom_sha<<16 | niter_soil<<24
  D_FLAGS = 62,  // uint32 bits
  // accumulated over every step since the caller last zeroed them (mm); read-modify-write (driver duty, not func:)
  D_ACCPRCP = 63, D_ACCECAN, D_ACCETRAN, D_ACCEDIR, D_SFCRNOFF, D_UGDRNOFF,
  NDIAG = 69
};

// packed parameter tables (see noahmp_b200/tables.py: pack_tables)
enum GenSlot {
  G_CSOIL = 0, G_ZBOT, G_CZIL, G_DKREF, G_KDTREF, G_FRZK, G_TIMEAN, G_FSATMAX, G_MLTFCT, G_Z0SNO, G_SSI,
  G_SWEMAX, G_ALBICE /*2*/, G_ALBLAKE = 14 /*2*/, G_OMEGAS = 16 /*2*/, G_BETADS = 18, G_BETAIS, G_EMSSOIL,
  G_EMSLAKE, G_NSLOPE = 22, G_SLOPE = 23 /*30*/, NGEN = 53
};
enum SoilSlot { P_BEXP = 0, P_SMCMAX, P_SMCREF, P_SMCWLT, P_PSISAT, P_DKSAT, P_DWSAT, P_QUARTZ, P_KDT, P_FRZX, NSOILP = 10 };
enum ColorSlot { K_ALBSAT_VIS = 0, K_ALBSAT_NIR, K_ALBDRY_VIS, K_ALBDRY_NIR, NCOLORP = 4 };
enum VegSlot {
  V_XL = 0, V_RHOL /*2*/, V_RHOS = 3, V_TAUL = 5, V_TAUS = 7, V_CANWMXP = 9, V_DLEAF, V_Z0MVT, V_HVT, V_HVB,
  V_DEN, V_RCROWN, V_CWPVT, V_LAI12M = 17 /*12*/, V_SAI12M = 29 /*12*/, V_SLA = 41, V_DILEFC, V_DILEFW,
  V_FRAGR, V_LTOVRC, V_WRRAT, V_WDPOOL, V_TDLEF, V_RGL = 49, V_HS, V_KC25, V_AKC, V_KO25, V_AKO, V_VCMX25,
  V_AVCMX, V_BP, V_RSMAX, V_RSMIN, V_MP, V_QE25, V_AQE, V_RMF25, V_RMS25, V_RMR25, V_FOLNMX, V_TOPT,
  V_TMIN, V_ARM, V_MRP, NVEGF = 71
};
enum VegISlot { VI_NROOT = 0, VI_C3C4, NVEGI = 2 };

// per-run scalars (group G of SURVEY App. A)
struct RunParams {
  double julian0;     // JULIAN of the first step; step s uses (float)(julian0 + s*dt/86400)
  float  dt;          // DT [s]
  float  zsoil[4];    // ZSOIL [m], negative
  float  zref;        // reference height above the displacement/canopy height [m] (ZLVL input)
  float  co2_frac;    // CO2AIR = co2_frac * SFCPRS   (355e-6)
  float  o2_frac;     // O2AIR  = o2_frac  * SFCPRS   (0.209)
  float  foln;        // FOLN [%]
  int    yearlen;     // YEARLEN [days]
};

}  // namespace noahmp_oracle
