// TEST INFRASTRUCTURE ONLY (see common.h).
// tsnosoi, hrt, hstep, rosr12, phasechange, frh2o : core/module_noahmp_func.f90:3987-4598
#include "physics.h"

namespace noahmp_oracle {

// func:3987-4096 (the energy check after the unconditional return at func:4071 is dead code)
void tsnosoi(const Ctx& cx, int ISNOW, float TBOT, const A7& ZSNSO, float SSOIL, const A7& DF,
             const A7& HCPCT, float ZBOT, float DT, float snowh, A7& STC) {
  A7 AI, BI, CI, RHSTS, PHI;
  for (int i = 0; i < 7; ++i) { AI.v[i] = 0.0f; BI.v[i] = 0.0f; CI.v[i] = 0.0f; RHSTS.v[i] = 0.0f; PHI.v[i] = 0.0f; }
  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) PHI(IZ) = 0.0f;   // func:4035
  float ZBOTSNO = ZBOT - snowh;                                 // func:4039
  float EFLXB = 0.0f;
  hrt(cx, ISNOW, ZSNSO, STC, TBOT, ZBOTSNO, DF, HCPCT, SSOIL, PHI, AI, BI, CI, RHSTS, EFLXB);  // func:4049-4053
  hstep(ISNOW, DT, AI, BI, CI, RHSTS, STC);                     // func:4055-4057
}

// func:4099-4188
void hrt(const Ctx& cx, int ISNOW, const A7& ZSNSO, const A7& STC, float TBOT, float ZBOT,
         const A7& DF, const A7& HCPCT, float SSOIL, const A7& PHI, A7& AI, A7& BI, A7& CI, A7& RHSTS,
         float& BOTFLX) {
  A7 DDZ, DENOM, DTSDZ, EFLUX;
  for (int i = 0; i < 7; ++i) { DDZ.v[i] = 0.0f; DENOM.v[i] = 0.0f; DTSDZ.v[i] = 0.0f; EFLUX.v[i] = 0.0f; }
  float TEMP1;
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {                    // func:4139-4164
    if (K == ISNOW + 1) {
      DENOM(K) = -ZSNSO(K) * HCPCT(K);
      TEMP1 = -ZSNSO(K + 1);
      DDZ(K) = 2.0f / TEMP1;
      DTSDZ(K) = 2.0f * (STC(K) - STC(K + 1)) / TEMP1;
      EFLUX(K) = DF(K) * DTSDZ(K) - SSOIL - PHI(K);
    } else if (K < NSOIL) {
      DENOM(K) = (ZSNSO(K - 1) - ZSNSO(K)) * HCPCT(K);
      TEMP1 = ZSNSO(K - 1) - ZSNSO(K + 1);
      DDZ(K) = 2.0f / TEMP1;
      DTSDZ(K) = 2.0f * (STC(K) - STC(K + 1)) / TEMP1;
      EFLUX(K) = (DF(K) * DTSDZ(K) - DF(K - 1) * DTSDZ(K - 1)) - PHI(K);
    } else if (K == NSOIL) {
      DENOM(K) = (ZSNSO(K - 1) - ZSNSO(K)) * HCPCT(K);
      TEMP1 = ZSNSO(K - 1) - ZSNSO(K);
      if (cx.opt.opt_tbot == 1) BOTFLX = 0.0f;
      if (cx.opt.opt_tbot == 2) {
        DTSDZ(K) = (STC(K) - TBOT) / (0.5f * (ZSNSO(K - 1) + ZSNSO(K)) - ZBOT);
        BOTFLX = -DF(K) * DTSDZ(K);
      }
      EFLUX(K) = (-BOTFLX - DF(K - 1) * DTSDZ(K - 1)) - PHI(K);
    }
  }

  for (int K = ISNOW + 1; K <= NSOIL; ++K) {                    // func:4166-4186
    if (K == ISNOW + 1) {
      AI(K) = 0.0f;
      CI(K) = -DF(K) * DDZ(K) / DENOM(K);
      if (cx.opt.opt_stc == 1) BI(K) = -CI(K);
      if (cx.opt.opt_stc == 2) BI(K) = -CI(K) + DF(K) / (0.5f * ZSNSO(K) * ZSNSO(K) * HCPCT(K));
    } else if (K < NSOIL) {
      AI(K) = -DF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = -DF(K) * DDZ(K) / DENOM(K);
      BI(K) = -(AI(K) + CI(K));
    } else if (K == NSOIL) {
      AI(K) = -DF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = 0.0f;
      BI(K) = -(AI(K) + CI(K));
    }
    RHSTS(K) = EFLUX(K) / (-DENOM(K));
  }
}

// func:4190-4237
void hstep(int ISNOW, float DT, A7& AI, A7& BI, A7& CI, A7& RHSTS, A7& STC) {
  A7 RHSTSIN, CIIN;
  for (int i = 0; i < 7; ++i) { RHSTSIN.v[i] = 0.0f; CIIN.v[i] = 0.0f; }
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {                    // func:4213-4218
    RHSTS(K) = RHSTS(K) * DT;
    AI(K) = AI(K) * DT;
    BI(K) = 1.0f + BI(K) * DT;
    CI(K) = CI(K) * DT;
  }
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {                    // func:4222-4225
    RHSTSIN(K) = RHSTS(K);
    CIIN(K) = CI(K);
  }
  // func:4229 -- arrays are (-NSNOW+1:NSOIL); pass base pointers and the NSNOW offset
  rosr12(CI.v, AI.v, BI.v, CIIN.v, RHSTSIN.v, RHSTS.v, ISNOW + 1, NSOIL, NSNOW);
  for (int K = ISNOW + 1; K <= NSOIL; ++K) STC(K) = STC(K) + CI(K);  // func:4233-4235
}

// func:4240-4288.  Arrays are dimensioned (-NSNOW_+1:NSOIL_): element k lives at [k + NSNOW_ - 1].
void rosr12(float* P, const float* A, const float* B, float* C, const float* D, float* DELTA, int NTOP,
            int NSOIL_, int NSNOW_) {
  const int off = NSNOW_ - 1;
#define IX(k) ((k) + off)
  C[IX(NSOIL_)] = 0.0f;                                         // func:4268
  P[IX(NTOP)] = -C[IX(NTOP)] / B[IX(NTOP)];                     // func:4269
  DELTA[IX(NTOP)] = D[IX(NTOP)] / B[IX(NTOP)];                  // func:4272
  for (int K = NTOP + 1; K <= NSOIL_; ++K) {                    // func:4275-4278
    P[IX(K)] = -C[IX(K)] * (1.0f / (B[IX(K)] + A[IX(K)] * P[IX(K - 1)]));
    DELTA[IX(K)] = (D[IX(K)] - A[IX(K)] * DELTA[IX(K - 1)]) * (1.0f / (B[IX(K)] + A[IX(K)] * P[IX(K - 1)]));
  }
  P[IX(NSOIL_)] = DELTA[IX(NSOIL_)];                            // func:4281
  for (int K = NTOP + 1; K <= NSOIL_; ++K) {                    // func:4284-4287
    int KK = NSOIL_ - K + (NTOP - 1) + 1;
    P[IX(KK)] = P[IX(KK)] * P[IX(KK + 1)] + DELTA[IX(KK)];
  }
#undef IX
}

// func:4291-4491
void phasechange(const Ctx& cx, int sltyp, int ISNOW, float DT, const A7& FACT, const A7& DZSNSO,
                 const A7& HCPCT, int IST, A7& STC, A3& SNICE, A3& SNLIQ, float& sneqv, float& snowh,
                 A4& SMC, A4& soilwat, float& QMELT, I7& IMELT, float& PONDING) {
  (void)HCPCT;
  const Tables& t = cx.tbl;
  A7 HM, XM, WMASS0, WICE0, WLIQ0, MICE, MLIQ, SUPERCOOL;
  for (int i = 0; i < 7; ++i) {
    HM.v[i] = 0.0f; XM.v[i] = 0.0f; WMASS0.v[i] = 0.0f; WICE0.v[i] = 0.0f; WLIQ0.v[i] = 0.0f;
    MICE.v[i] = 0.0f; MLIQ.v[i] = 0.0f; IMELT.v[i] = 0;
  }
  QMELT = 0.0f;                                                 // func:4346-4348
  PONDING = 0.0f;
  float XMF = 0.0f;

  for (int J = -NSNOW + 1; J <= NSOIL; ++J) SUPERCOOL(J) = 0.0f;  // func:4350-4352

  for (int J = ISNOW + 1; J <= 0; ++J) {                        // func:4354-4357
    MICE(J) = SNICE(J);
    MLIQ(J) = SNLIQ(J);
  }
  for (int J = 1; J <= NSOIL; ++J) {                            // func:4359-4362
    MLIQ(J) = soilwat(J) * DZSNSO(J) * 1000.0f;
    MICE(J) = (SMC(J) - soilwat(J)) * DZSNSO(J) * 1000.0f;
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {                    // func:4364-4371
    IMELT(J) = 0;
    HM(J) = 0.0f;
    XM(J) = 0.0f;
    WICE0(J) = MICE(J);
    WLIQ0(J) = MLIQ(J);
    WMASS0(J) = MICE(J) + MLIQ(J);
  }

  if (IST == 1) {                                               // func:4373-4387
    for (int J = 1; J <= NSOIL; ++J) {
      if (cx.opt.opt_frz == 1) {
        if (STC(J) < TFRZ) {
          float SMP = HFUS * (TFRZ - STC(J)) / (GRAV * STC(J));
          SUPERCOOL(J) = t.LK_SMCMAX[sltyp] * powf(SMP / t.LK_PSISAT[sltyp], -1.0f / t.LK_BEXP[sltyp]);
          SUPERCOOL(J) = SUPERCOOL(J) * DZSNSO(J) * 1000.0f;
        }
      }
      if (cx.opt.opt_frz == 2) {
        frh2o(cx, sltyp, SUPERCOOL(J), STC(J), SMC(J), soilwat(J));
        SUPERCOOL(J) = SUPERCOOL(J) * DZSNSO(J) * 1000.0f;
      }
    }
  }

  for (int J = ISNOW + 1; J <= NSOIL; ++J) {                    // func:4389-4403
    if (MICE(J) > 0.0f && STC(J) >= TFRZ) IMELT(J) = 1;
    if (MLIQ(J) > SUPERCOOL(J) && STC(J) < TFRZ) IMELT(J) = 2;
    if (ISNOW == 0 && sneqv > 0.0f && J == 1) {
      if (STC(J) >= TFRZ) IMELT(J) = 1;
    }
  }

  for (int J = ISNOW + 1; J <= NSOIL; ++J) {                    // func:4406-4421
    if (IMELT(J) > 0) {
      HM(J) = (STC(J) - TFRZ) / FACT(J);
      STC(J) = TFRZ;
    }
    if (IMELT(J) == 1 && HM(J) < 0.0f) {
      HM(J) = 0.0f;
      IMELT(J) = 0;
    }
    if (IMELT(J) == 2 && HM(J) > 0.0f) {
      HM(J) = 0.0f;
      IMELT(J) = 0;
    }
    XM(J) = HM(J) * DT / HFUS;
  }

  if (ISNOW == 0 && sneqv > 0.0f && XM(1) > 0.0f) {             // func:4424-4440
    float TEMP1 = sneqv;
    sneqv = fmax2(0.0f, TEMP1 - XM(1));
    float PROPOR = sneqv / TEMP1;
    snowh = fmax2(0.0f, PROPOR * snowh);
    float HEATR = HM(1) - HFUS * (TEMP1 - sneqv) / DT;
    if (HEATR > 0.0f) {
      XM(1) = HEATR * DT / HFUS;
      HM(1) = HEATR;
    } else {
      XM(1) = 0.0f;
      HM(1) = 0.0f;
    }
    QMELT = fmax2(0.0f, (TEMP1 - sneqv)) / DT;
    XMF = HFUS * QMELT;
    PONDING = TEMP1 - sneqv;
  }

  for (int J = ISNOW + 1; J <= NSOIL; ++J) {                    // func:4443-4479
    if (IMELT(J) > 0 && fabsf(HM(J)) > 0.0f) {
      float HEATR = 0.0f;
      if (XM(J) > 0.0f) {
        MICE(J) = fmax2(0.0f, WICE0(J) - XM(J));
        HEATR = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
      } else if (XM(J) < 0.0f) {
        if (J <= 0) {
          MICE(J) = fmin2(WMASS0(J), WICE0(J) - XM(J));
        } else {
          if (WMASS0(J) < SUPERCOOL(J)) {
            MICE(J) = 0.0f;
          } else {
            MICE(J) = fmin2(WMASS0(J) - SUPERCOOL(J), WICE0(J) - XM(J));
            MICE(J) = fmax2(MICE(J), 0.0f);
          }
        }
        HEATR = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
      }

      MLIQ(J) = fmax2(0.0f, WMASS0(J) - MICE(J));               // func:4464

      if (fabsf(HEATR) > 0.0f) {                                // func:4466-4471
        STC(J) = STC(J) + FACT(J) * HEATR;
        if (J <= 0) {
          if (MLIQ(J) * MICE(J) > 0.0f) STC(J) = TFRZ;
        }
      }

      XMF = XMF + HFUS * (WICE0(J) - MICE(J)) / DT;             // func:4473

      if (J < 1) QMELT = QMELT + fmax2(0.0f, (WICE0(J) - MICE(J))) / DT;  // func:4475-4477
    }
  }
  (void)XMF;

  for (int J = ISNOW + 1; J <= 0; ++J) {                        // func:4481-4484
    SNLIQ(J) = MLIQ(J);
    SNICE(J) = MICE(J);
  }
  for (int J = 1; J <= NSOIL; ++J) {                            // func:4486-4489
    soilwat(J) = MLIQ(J) / (1000.0f * DZSNSO(J));
    SMC(J) = (MLIQ(J) + MICE(J)) / (1000.0f * DZSNSO(J));
  }
}

// func:4494-4598
void frh2o(const Ctx& cx, int sltyp, float& FREE, float TKELV, float SMC, float soilwat) {
  const Tables& t = cx.tbl;
  const float CK = 8.0f, BLIM = 5.5f, ERROR = 0.005f;
  float BX = t.LK_BEXP[sltyp];                                  // func:4539-4546
  if (t.LK_BEXP[sltyp] > BLIM) BX = BLIM;
  int NLOG = 0;
  int KCOUNT = 0;
  if (TKELV > (TFRZ - 1.0E-3f)) {                               // func:4547-4548
    FREE = SMC;
  } else {
    if (CK != 0.0f) {                                           // func:4552-4583
      float SWL = SMC - soilwat;
      if (SWL > (SMC - 0.02f)) SWL = SMC - 0.02f;
      if (SWL < 0.0f) SWL = 0.0f;
      while ((NLOG < 10) && (KCOUNT == 0)) {
        NLOG = NLOG + 1;
        float DF = logf((t.LK_PSISAT[sltyp] * GRAV / HFUS) * (powf(1.0f + CK * SWL, 2.0f)) *
                        powf(t.LK_SMCMAX[sltyp] / (SMC - SWL), BX)) -
                   logf(-(TKELV - TFRZ) / TKELV);
        float DENOM = 2.0f * CK / (1.0f + CK * SWL) + BX / (SMC - SWL);
        float SWLK = SWL - DF / DENOM;
        if (SWLK > (SMC - 0.02f)) SWLK = SMC - 0.02f;
        if (SWLK < 0.0f) SWLK = 0.0f;
        float DSWL = fabsf(SWLK - SWL);
        SWL = SWLK;
        if (DSWL <= ERROR) KCOUNT = KCOUNT + 1;
      }
      FREE = SMC - SWL;
    }
    if (KCOUNT == 0) {                                          // func:4588-4595
      float FK = (powf((HFUS / (GRAV * (-t.LK_PSISAT[sltyp]))) * ((TKELV - TFRZ) / TKELV), -1.0f / BX)) *
                 t.LK_SMCMAX[sltyp];
      if (FK < 0.02f) FK = 0.02f;
      FREE = fmin2(FK, SMC);
    }
  }
}

}  // namespace noahmp_oracle
