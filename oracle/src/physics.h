// TEST INFRASTRUCTURE ONLY (see common.h).  Internal prototypes, one per reference subroutine
// of core/module_noahmp_func.f90; argument order follows the Fortran where practical, minus the
// never-used ILOC/JLOC/DX/DZ8W/QC/PBLH/IZ0TLND pass-throughs.
#pragma once
#include "common.h"

namespace noahmp_oracle {

// func:479-531
void atm(float SFCPRS, float SFCTMP, float Q2, float PRCP, float SOLDN, float COSZ, float& THAIR,
         float& QAIR, float& EAIR, float& RHOAIR, float& QPRECC, float& QPRECL, A2& SOLAD, A2& SOLAI,
         float& SWDOWN);
// func:534-630
void phenology(const Ctx& cx, int lutyp, float snowh, float TV, float LAT, int YEARLEN, float JULIAN,
               float& LAI, float& SAI, float TROOT, float& HTOP, float& elai, float& esai, float& IGS);

struct EnergyOut {
  I7 IMELT;
  A3 SNICEV, SNLIQV, epore;
  float T2M, QMELT, TAUX, TAUY, TS, LATHEAV, LATHEAG, FSRV, FSRG, Q1, Q2E;
  A4 BTRANI;
  float BTRAN;
  bool frozen_canopy, frozen_ground;
};
// func:735-1338
void energy(const Ctx& cx, ColumnIO& c, float RHOAIR, float QAIR, float THAIR, float EAIR,
            const A2& SOLAD, const A2& SOLAI, float IGS, float HTOP, float elai, float esai,
            A7& DZSNSO, EnergyOut& o);

// func:1341-1445
void thermoprop(const Ctx& cx, int sltyp, int ISNOW, int IST, const A7& DZSNSO, float DT, float snowh,
                const A3& SNICE, const A3& SNLIQ, float CSOIL, const A4& SMC, const A4& soilwat,
                const A7& STC, int lutyp, A7& DF, A7& HCPCT, A3& SNICEV, A3& SNLIQV, A3& epore,
                A7& FACT);
// func:1448-1497
void csnow(int ISNOW, const A3& SNICE, const A3& SNLIQ, const A7& DZSNSO, A3& TKSNO, A3& CVSNO,
           A3& SNICEV, A3& SNLIQV, A3& epore);

// func:1598-1714
void radiation(const Ctx& cx, int lutyp, int IST, int ISC, int ICE, float SNEQVO, float sneqv,
               float DT, float COSZ, float snowh, float TG, float TV, float FSNO, float QSNOW,
               float FWET, float elai, float esai, const A4& SMC, const A2& SOLAD, const A2& SOLAI,
               float fveg, float& ALBOLD, float& TAUSS, float& FSUN, float& LAISUN, float& LAISHA,
               float& PARSUN, float& PARSHA, float& SAV, float& SAG, float& FSR, float& FSA,
               float& FSRV, float& FSRG, float& BGAP, float& WGAP);
// func:1717-1887
void albedo(const Ctx& cx, int lutyp, int IST, int ISC, int ICE, float DT, float COSZ, float& FAGE,
            float elai, float esai, float TG, float TV, float snowh, float FSNO, float FWET,
            const A4& SMC, float SNEQVO, float sneqv, float QSNOW, float fveg, float& ALBOLD,
            float& TAUSS, A2& ALBGRD, A2& ALBGRI, A2& ALBD, A2& ALBI, A2& FABD, A2& FABI, A2& FTDD,
            A2& FTID, A2& FTII, float& FSUN, A2& FREVI, A2& FREVD, A2& FREGD, A2& FREGI, float& BGAP,
            float& WGAP);
// func:1890-2005
void surrad(float MPE_, float FSUN, float FSHA, float elai, float VAI, float LAISUN, float LAISHA,
            const A2& SOLAD, const A2& SOLAI, const A2& FABD, const A2& FABI, const A2& FTDD,
            const A2& FTID, const A2& FTII, const A2& ALBGRD, const A2& ALBGRI, const A2& ALBD,
            const A2& ALBI, float& PARSUN, float& PARSHA, float& SAV, float& SAG, float& FSA,
            float& FSR, const A2& FREVI, const A2& FREVD, const A2& FREGD, const A2& FREGI,
            float& FSRV, float& FSRG);
// func:2008-2054
void snowage(const Ctx& cx, float DT, float TG, float SNEQVO, float sneqv, float& TAUSS, float& FAGE);
// func:2057-2102
void snowalb_bats(float FSNO, float COSZ, float FAGE, A2& ALBSND, A2& ALBSNI);
// func:2105-2151
void snowalb_class(const Ctx& cx, float QSNOW, float DT, float& ALB, float ALBOLD, A2& ALBSND,
                   A2& ALBSNI);
// func:2154-2212
void groundalb(const Ctx& cx, int ICE, int IST, int ISC, float FSNO, const A4& SMC, const A2& ALBSND,
               const A2& ALBSNI, float COSZ, float TG, A2& ALBGRD, A2& ALBGRI);

struct VegeFluxIO;  // flux.cpp
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
               float PSFC, float& Q2V, float& CAH2, float& CHLEAF, float& CHUC);
// func:2967-3257
void bare_flux(const Ctx& cx, ColumnIO& c, int lutyp, float DT, int ISNOW, float SAG, float LWDN,
               float UR, float UU, float VV, float SFCTMP, float THAIR, float QAIR, float EAIR,
               float RHOAIR, float snowh, const A7& DZSNSO, float ZLVL, float ZPD, float Z0M,
               float EMG, const A7& STC, const A7& DF, float RSURF, float LATHEA, float GAMMA,
               float RHSUR, float& TGB, float& CM, float& CH, float& TAUXB, float& TAUYB, float& IRB,
               float& SHB, float& EVB, float& GHB, float& T2MB, float& QSFC, float PSFC, float SFCPRS,
               float& Q2B, float& EHB2);
// func:3260-3350
void ragrb(const Ctx& cx, int iter, float VAI, float RHOAIR, float HG, float TAH, float ZPD,
           float Z0MG, float Z0HG, float HCAN, float UC, float Z0H, float FV, float CWP, int lutyp,
           float MPE_, float TV, float& MOZG, float& FHG, float& RAMG, float& RAHG, float& RAWG,
           float& RB);
// func:3353-3508
void sfcdif1(int iter, float SFCTMP, float RHOAIR, float H, float QAIR, float ZLVL, float ZPD,
             float Z0M, float Z0H, float UR, float MPE_, float& MOZ, int& MOZSGN, float& FM, float& FH,
             float& FM2, float& FH2, float& CM, float& CH, float& FV, float& CH2, uint32_t& flags);
// func:3511-3689
void sfcdif2(int iter, float Z0, float THZ0, float THLM, float SFCSPD, float CZIL, float ZLM,
             float& AKMS, float& AKHS, float& RLMO, float& WSTAR2, float& USTAR);
// func:3890-3955, 3958-3984
void canres(const Ctx& cx, int lutyp, float sfcprs, float tv, float par, float eah, float btran,
            float& rs, float& psn);
void calhum(float SFCTMP, float SFCPRS, float& Q2SAT, float& DQSDT2);

// func:3987-4096
void tsnosoi(const Ctx& cx, int ISNOW, float TBOT, const A7& ZSNSO, float SSOIL, const A7& DF,
             const A7& HCPCT, float ZBOT, float DT, float snowh, A7& STC);
// func:4099-4188
void hrt(const Ctx& cx, int ISNOW, const A7& ZSNSO, const A7& STC, float TBOT, float ZBOT,
         const A7& DF, const A7& HCPCT, float SSOIL, const A7& PHI, A7& AI, A7& BI, A7& CI, A7& RHSTS,
         float& BOTFLX);
// func:4190-4237
void hstep(int ISNOW, float DT, A7& AI, A7& BI, A7& CI, A7& RHSTS, A7& STC);
// func:4291-4491
void phasechange(const Ctx& cx, int sltyp, int ISNOW, float DT, const A7& FACT, const A7& DZSNSO,
                 const A7& HCPCT, int IST, A7& STC, A3& SNICE, A3& SNLIQ, float& sneqv, float& snowh,
                 A4& SMC, A4& soilwat, float& QMELT, I7& IMELT, float& PONDING);

// func:4601-4804
void water(const Ctx& cx, ColumnIO& c, const I7& IMELT, float QPRECC, float QPRECL, float elai,
           float esai, float QVAP, float QDEW, const A4& BTRANI, float PONDING, bool frozen_canopy,
           bool frozen_ground, A4& soilice, A7& DZSNSO, float& CMC, float& QIN, float& QDIS);
// func:4807-5046
void canwater(const Ctx& cx, int lutyp, float DT, float SFCTMP, float UU, float VV, float FCEV,
              float FCTR, float QPRECC, float QPRECL, float elai, float esai, int IST, float TG,
              float fveg, bool FROZEN_CANOPY, float& CANLIQ, float& CANICE, float& TV, float& CMC,
              float& ECAN, float& ETRAN, float& QRAIN, float& QSNOW, float& snowhin, float& FWET,
              float& FPICE);
// func:5049-5174
void snowwater(const Ctx& cx, float DT, const A4& ZSOIL, const I7& IMELT, float SFCTMP, float snowhin,
               float QSNOW, float QSNFRO, float QSNSUB, float QRAIN, const A3& FICEOLD, int& ISNOW,
               float& snowh, float& sneqv, A3& SNICE, A3& SNLIQ, A4& soilwat, A4& soilice, A7& STC,
               A7& ZSNSO, A7& DZSNSO, float& QSNBOT, float& SNOFLOW, float& PONDING1,
               float& PONDING2);
// func:5177-5233
void snowfall(float DT, float QSNOW, float snowhin, float SFCTMP, int& ISNOW, float& snowh,
              A7& DZSNSO, A7& STC, A3& SNICE, A3& SNLIQ, float& sneqv);
// func:5236-5413
void combine(int& ISNOW, A4& soilwat, A7& STC, A3& SNICE, A3& SNLIQ, A7& DZSNSO, A4& soilice,
             float& snowh, float& sneqv, float& PONDING1, float& PONDING2);
// func:5416-5533
void divide(int& ISNOW, A7& STC, A3& SNICE, A3& SNLIQ, A7& DZSNSO);
// func:5580-5677
void compact(float DT, const A7& STC, const A3& SNICE, const A3& SNLIQ, const I7& IMELT,
             const A3& FICEOLD, int ISNOW, A7& DZSNSO);
// func:5680-5819
void snowh2o(const Ctx& cx, float DT, float QSNFRO, float QSNSUB, float QRAIN, int& ISNOW,
             A7& DZSNSO, float& snowh, float& sneqv, A3& SNICE, A3& SNLIQ, A4& soilwat, A4& soilice,
             A7& STC, float& QSNBOT, float& PONDING1, float& PONDING2);
// func:5822-6048
void soilh2o(const Ctx& cx, int sltyp, int lutyp, float DT, const A4& ZSOIL, const A7& DZSNSO,
             int slptyp, float qinsrf, float QSEVA, const A4& ETRANI, const A4& soilice, A4& soilwat,
             A4& SMC, float& ZWT, float& runsrf, float& QDRAIN, float& runsub, A4& WCND,
             float& FCRMAX, int& niter_out);
// func:6051-6100
void zwteq(const Ctx& cx, int sltyp, const A4& ZSOIL, const A7& DZSNSO, const A4& soilwat, float& ZWT);
// func:6103-6196
void infil(const Ctx& cx, int sltyp, float DT, const A4& ZSOIL, const A4& soilwat, const A4& soilice,
           float SICEMAX, float qinsrf, float& qinfil, float& runsrf);
// func:6199-6305
void srt(const Ctx& cx, int sltyp, float DT, const A4& ZSOIL, int slptyp, float qinfil,
         const A4& ETRANI, float QSEVA, const A4& soilwat, const A4& SMC, float ZWT, const A4& FCR,
         float SICEMAX, float FCRMAX, A4& RHSTT, A4& AI, A4& BI, A4& CI, float& QDRAIN, A4& WCND);
// func:6308-6383
void sstep(const Ctx& cx, int sltyp, float DT, const A7& DZSNSO, const A4& soilice, A4& soilwat,
           A4& SMC, A4& AI, A4& BI, A4& CI, A4& RHSTT, float& WPLUS);
// func:6386-6455
void wdfcnd1(const Ctx& cx, int sltyp, float& WDF, float& WCND, float SMC, float FCR);
void wdfcnd2(const Ctx& cx, int sltyp, float& WDF, float& WCND, float SMC, float soilice);
// func:6458-6639
void groundwater(const Ctx& cx, int sltyp, float DT, const A4& ZSOIL, const A4& soilice,
                 const A4& WCND, float FCRMAX, A4& soilwat, float& ZWT, float& WA, float& WT,
                 float& QIN, float& QDIS);
// func:6642-6751, 6754-7025
void carbon(const Ctx& cx, int lutyp, float DT, const A4& ZSOIL, const A7& DZSNSO, const A7& STC,
            const A4& SMC, float TV, float TG, float PSN, float FOLN, float SMCMAX, float BTRAN,
            float APAR, float fveg, float IGS, float TROOT, int IST, float LAT, float& LFMASS,
            float& RTMASS, float& STMASS, float& WOOD, float& STBLCP, float& FASTCP, float& GPP,
            float& NPP, float& NEE, float& AUTORS, float& HETERS, float& TOTSC, float& TOTLB,
            float& XLAI, float& XSAI);
void co2flux(const Ctx& cx, int lutyp, float IGS, float DT, const A7& DZSNSO, const A7& STC,
             float PSN, float TROOT, float TV, float WROOT, float WSTRES, float FOLN, float LAPM,
             float LAT, float fveg, float& XLAI, float& XSAI, float& LFMASS, float& RTMASS,
             float& STMASS, float& FASTCP, float& STBLCP, float& WOOD, float& GPP, float& NPP,
             float& NEE, float& AUTORS, float& HETERS, float& TOTSC, float& TOTLB);
// func:633-732
void error(ColumnIO& c, float SWDOWN, float BEG_WB, const A7& DZSNSO);

}  // namespace noahmp_oracle
