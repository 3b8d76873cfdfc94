// TEST INFRASTRUCTURE ONLY (see common.h).
// C ABI of the CPU oracle: same packed tables / SoA planes as the product library so a test can
// hand both the same buffers.  OpenMP over columns (columns are independent: func:66 is a
// one-column routine and nothing in core/ loops over cells).
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "layout.h"
#include "physics.h"

using namespace noahmp_oracle;

namespace {

void unpack_tables(Tables& t, const float* gen, int ngen, const float* soil, int nsoiltyp,
                   const float* color, int ncolor, const float* veg_f, const int* veg_i, int nvegtyp,
                   const int* flags) {
  (void)ngen;
  std::memset(&t, 0, sizeof(t));
  t.KK_CSOIL = gen[G_CSOIL]; t.KK_ZBOT = gen[G_ZBOT]; t.KK_CZIL = gen[G_CZIL]; t.KK_DKREF = gen[G_DKREF];
  t.KK_KDTREF = gen[G_KDTREF]; t.KK_FRZK = gen[G_FRZK]; t.KK_TIMEAN = gen[G_TIMEAN];
  t.KK_FSATMAX = gen[G_FSATMAX]; t.KK_MLTFCT = gen[G_MLTFCT]; t.KK_Z0SNO = gen[G_Z0SNO];
  t.KK_SSI = gen[G_SSI]; t.KK_SWEMAX = gen[G_SWEMAX];
  for (int b = 1; b <= 2; ++b) {
    t.KK_ALBICE[b] = gen[G_ALBICE + b - 1];
    t.KK_ALBLAKE[b] = gen[G_ALBLAKE + b - 1];
    t.KK_OMEGAS[b] = gen[G_OMEGAS + b - 1];
  }
  t.KK_BETADS = gen[G_BETADS]; t.KK_BETAIS = gen[G_BETAIS]; t.KK_EMSSOIL = gen[G_EMSSOIL];
  t.KK_EMSLAKE = gen[G_EMSLAKE];
  for (int i = 1; i <= MSLOPETYP; ++i) t.LK_SLOPE[i] = gen[G_SLOPE + i - 1];
  t.nsltyp = nsoiltyp;
  for (int i = 1; i <= nsoiltyp && i <= MSLTYP; ++i) {
    const float* r = soil + (i - 1) * NSOILP;
    t.LK_BEXP[i] = r[P_BEXP]; t.LK_SMCMAX[i] = r[P_SMCMAX]; t.LK_SMCREF[i] = r[P_SMCREF];
    t.LK_SMCWLT[i] = r[P_SMCWLT]; t.LK_PSISAT[i] = r[P_PSISAT]; t.LK_DKSAT[i] = r[P_DKSAT];
    t.LK_DWSAT[i] = r[P_DWSAT]; t.LK_QUARTZ[i] = r[P_QUARTZ]; t.LK_KDT[i] = r[P_KDT];
    t.LK_FRZX[i] = r[P_FRZX];
  }
  t.nsoilcol = ncolor;
  for (int i = 1; i <= ncolor && i <= MSLCOL; ++i) {
    const float* r = color + (i - 1) * NCOLORP;
    t.LK_ALBSAT[1][i] = r[K_ALBSAT_VIS]; t.LK_ALBSAT[2][i] = r[K_ALBSAT_NIR];
    t.LK_ALBDRY[1][i] = r[K_ALBDRY_VIS]; t.LK_ALBDRY[2][i] = r[K_ALBDRY_NIR];
  }
  t.nlutyp = nvegtyp;
  t.ISURBAN = flags[0]; t.ISWATER = flags[1]; t.ISBARREN = flags[2]; t.ISICE = flags[3]; t.ISEGBLF = flags[4];
  for (int i = 1; i <= nvegtyp && i <= MLUTYP; ++i) {
    const float* r = veg_f + (i - 1) * NVEGF;
    const int* ri = veg_i + (i - 1) * NVEGI;
    t.LK_XL[i] = r[V_XL];
    for (int b = 1; b <= 2; ++b) {
      t.LK_RHOL[b][i] = r[V_RHOL + b - 1]; t.LK_RHOS[b][i] = r[V_RHOS + b - 1];
      t.LK_TAUL[b][i] = r[V_TAUL + b - 1]; t.LK_TAUS[b][i] = r[V_TAUS + b - 1];
    }
    t.LK_NROOT[i] = ri[VI_NROOT]; t.LK_C3C4[i] = ri[VI_C3C4];
    t.LK_CANWMXP[i] = r[V_CANWMXP]; t.LK_DLEAF[i] = r[V_DLEAF]; t.LK_Z0MVT[i] = r[V_Z0MVT];
    t.LK_HVT[i] = r[V_HVT]; t.LK_HVB[i] = r[V_HVB]; t.LK_DEN[i] = r[V_DEN]; t.LK_RCROWN[i] = r[V_RCROWN];
    t.LK_CWPVT[i] = r[V_CWPVT];
    for (int m = 1; m <= 12; ++m) {
      t.LK_LAI12M[m][i] = r[V_LAI12M + m - 1];
      t.LK_SAI12M[m][i] = r[V_SAI12M + m - 1];
    }
    t.LK_SLA[i] = r[V_SLA]; t.LK_DILEFC[i] = r[V_DILEFC]; t.LK_DILEFW[i] = r[V_DILEFW];
    t.LK_FRAGR[i] = r[V_FRAGR]; t.LK_LTOVRC[i] = r[V_LTOVRC]; t.LK_WRRAT[i] = r[V_WRRAT];
    t.LK_WDPOOL[i] = r[V_WDPOOL]; t.LK_TDLEF[i] = r[V_TDLEF];
    t.LK_RGL[i] = r[V_RGL]; t.LK_HS[i] = r[V_HS]; t.LK_KC25[i] = r[V_KC25]; t.LK_AKC[i] = r[V_AKC];
    t.LK_KO25[i] = r[V_KO25]; t.LK_AKO[i] = r[V_AKO]; t.LK_VCMX25[i] = r[V_VCMX25];
    t.LK_AVCMX[i] = r[V_AVCMX]; t.LK_BP[i] = r[V_BP]; t.LK_RSMAX[i] = r[V_RSMAX];
    t.LK_RSMIN[i] = r[V_RSMIN]; t.LK_MP[i] = r[V_MP]; t.LK_QE25[i] = r[V_QE25]; t.LK_AQE[i] = r[V_AQE];
    t.LK_RMF25[i] = r[V_RMF25]; t.LK_RMS25[i] = r[V_RMS25]; t.LK_RMR25[i] = r[V_RMR25];
    t.LK_FOLNMX[i] = r[V_FOLNMX]; t.LK_TOPT[i] = r[V_TOPT]; t.LK_TMIN[i] = r[V_TMIN];
    t.LK_ARM[i] = r[V_ARM]; t.LK_MRP[i] = r[V_MRP];
  }
}

inline int f2i(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline float i2f(int i) { float f; std::memcpy(&f, &i, 4); return f; }

}  // namespace

extern "C" {

struct oracle_ctx {
  Ctx cx;
  bool have_tables = false, have_opts = false;
};

oracle_ctx* oracle_create() { return new oracle_ctx(); }
void oracle_destroy(oracle_ctx* c) { delete c; }

// order of core/module_noahmp_global.f90:77-80
int oracle_set_options(oracle_ctx* c, const int opts[12]) {
  Options& o = c->cx.opt;
  o.opt_veg = opts[0]; o.opt_crs = opts[1]; o.opt_btr = opts[2]; o.opt_run = opts[3]; o.opt_sfc = opts[4];
  o.opt_frz = opts[5]; o.opt_inf = opts[6]; o.opt_rad = opts[7]; o.opt_alb = opts[8]; o.opt_snf = opts[9];
  o.opt_tbot = opts[10]; o.opt_stc = opts[11];
  c->have_opts = true;
  return 0;
}

int oracle_set_tables(oracle_ctx* c, const float* gen, int ngen, const float* soil, int nsoiltyp,
                      const float* color, int ncolor, const float* veg_f, const int* veg_i, int nvegtyp,
                      const int* flags) {
  if (ngen != NGEN) return -1;
  unpack_tables(c->cx.tbl, gen, ngen, soil, nsoiltyp, color, ncolor, veg_f, veg_i, nvegtyp, flags);
  c->have_tables = true;
  return 0;
}

int oracle_layout(int what) {
  switch (what) {
    case 0: return NSTATE;
    case 1: return NSTATE_CORE;
    case 2: return NFORCING;
    case 3: return NSTATIC_F;
    case 4: return NDIAG;
    case 5: return NGEN;
    case 6: return NSOILP;
    case 7: return NCOLORP;
    case 8: return NVEGF;
    case 9: return NVEGI;
    case 10: return (int)sizeof(RunParams);
  }
  return -1;
}

// Advance ncol columns by nsteps.  state: [NSTATE][ncol] (in/out), forcing: [nsteps][NFORCING][ncol],
// static_f: [NSTATIC_F][ncol], static_i: [ncol] packed types, diag: [NDIAG][ncol] or NULL (values of
// the LAST step).  Returns the number of columns with a non-zero flag word on any step, <0 on misuse.
long long oracle_run(oracle_ctx* c, const RunParams* rp, long long ncol, int nsteps, float* state,
                     const float* forcing, const float* static_f, const int* static_i, float* diag,
                     int nthreads) {
  if (!c || !c->have_tables || !c->have_opts) return -1;
  const Ctx& cx = c->cx;
  long long nflag = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(static) reduction(+ : nflag)
  for (long long i = 0; i < ncol; ++i) {
    ColumnIO col;
    std::memset(&col, 0, sizeof(col));
    auto S = [&](int p) -> float& { return state[(size_t)p * ncol + i]; };
    // --- static
    col.DT = rp->dt;
    for (int k = 1; k <= 4; ++k) col.ZSOIL(k) = rp->zsoil[k - 1];
    col.FOLN = rp->foln;
    col.YEARLEN = rp->yearlen;
    col.ZLVL = rp->zref;
    col.LAT = static_f[(size_t)C_LAT * ncol + i];
    col.SHDFAC = static_f[(size_t)C_SHDFAC * ncol + i];
    col.TBOT = static_f[(size_t)C_TBOT * ncol + i];
    col.SHDMAX = static_f[(size_t)C_SHDMAX * ncol + i];
    const int ty = static_i[i];
    col.lutyp = ty & 0xff; col.sltyp = (ty >> 8) & 0xff; col.slptyp = (ty >> 16) & 0xff;
    col.ISC = (ty >> 24) & 0xf; col.IST = ((ty >> 28) & 1) + 1; col.ICE = (ty >> 29) & 1;
    // --- state in
    col.ALBOLD = S(S_ALBOLD); col.SNEQVO = S(S_SNEQVO); col.EAH = S(S_EAH); col.FWET = S(S_FWET);
    col.CANLIQ = S(S_CANLIQ); col.CANICE = S(S_CANICE); col.TV = S(S_TV); col.TG = S(S_TG);
    col.QSNOW = S(S_QSNOW); col.ISNOW = f2i(S(S_ISNOW)); col.snowh = S(S_SNOWH); col.sneqv = S(S_SNEQV);
    col.ZWT = S(S_ZWT); col.WA = S(S_WA); col.WT = S(S_WT); col.TAUSS = S(S_TAUSS);
    for (int k = 0; k < 7; ++k) { col.STC.v[k] = S(S_STC + k); col.ZSNSO.v[k] = S(S_ZSNSO + k); }
    for (int k = 0; k < 4; ++k) { col.soilwat.v[k] = S(S_SOILWAT + k); col.SMC.v[k] = S(S_SMC + k); }
    for (int k = 0; k < 3; ++k) { col.SNICE.v[k] = S(S_SNICE + k); col.SNLIQ.v[k] = S(S_SNLIQ + k); }
    col.QSFC = S(S_QSFC); col.TAH = S(S_TAH); col.CM = S(S_CM); col.CH = S(S_CH); col.WSLAKE = S(S_WSLAKE);
    col.LFMASS = S(S_LFMASS); col.RTMASS = S(S_RTMASS); col.STMASS = S(S_STMASS); col.WOOD = S(S_WOOD);
    col.STBLCP = S(S_STBLCP); col.FASTCP = S(S_FASTCP); col.LAI = S(S_LAI); col.SAI = S(S_SAI);

    uint32_t anyflag = 0;
    float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // accumulations between outputs (driver duty)
    if (diag) for (int q = 0; q < 6; ++q) acc[q] = diag[(size_t)(D_ACCPRCP + q) * ncol + i];
    for (int s = 0; s < nsteps; ++s) {
      const float* f = forcing + ((size_t)s * NFORCING) * ncol + i;
      col.SFCTMP = f[(size_t)F_SFCTMP * ncol]; col.SFCPRS = f[(size_t)F_SFCPRS * ncol];
      col.UU = f[(size_t)F_UU * ncol]; col.VV = f[(size_t)F_VV * ncol]; col.Q2 = f[(size_t)F_Q2 * ncol];
      col.SOLDN = f[(size_t)F_SOLDN * ncol]; col.LWDN = f[(size_t)F_LWDN * ncol];
      col.PRCP = f[(size_t)F_PRCP * ncol]; col.COSZ = f[(size_t)F_COSZ * ncol];
      col.PSFC = col.SFCPRS;                          // offline: lowest-model-level pressure = surface
      col.CO2AIR = rp->co2_frac * col.SFCPRS;
      col.O2AIR = rp->o2_frac * col.SFCPRS;
      {  // driver duty: day of year wrapped into [0, YEARLEN) (a launch may cross 31 December)
        const double jd = rp->julian0 + (double)s * (double)rp->dt / 86400.0, yl = (double)rp->yearlen;
        col.JULIAN = (float)(jd - yl * floor(jd / yl));
      }
      // FICEOLD: ice fraction of the snow layers at the start of the step (driver duty)
      for (int k = -2; k <= 0; ++k) {
        if (k >= col.ISNOW + 1) col.FICEOLD(k) = col.SNICE(k) / (col.SNICE(k) + col.SNLIQ(k));
        else col.FICEOLD(k) = 0.0f;
      }
      noahmp_sflx(cx, col);
      anyflag |= col.flags;
      acc[0] = acc[0] + col.PRCP * rp->dt; acc[1] = acc[1] + col.ECAN * rp->dt; acc[2] = acc[2] + col.ETRAN * rp->dt;
      acc[3] = acc[3] + col.EDIR * rp->dt; acc[4] = acc[4] + col.runsrf * rp->dt; acc[5] = acc[5] + col.runsub * rp->dt;
    }
    if (anyflag) nflag += 1;

    // --- state out
    S(S_ALBOLD) = col.ALBOLD; S(S_SNEQVO) = col.SNEQVO; S(S_EAH) = col.EAH; S(S_FWET) = col.FWET;
    S(S_CANLIQ) = col.CANLIQ; S(S_CANICE) = col.CANICE; S(S_TV) = col.TV; S(S_TG) = col.TG;
    S(S_QSNOW) = col.QSNOW; S(S_ISNOW) = i2f(col.ISNOW); S(S_SNOWH) = col.snowh; S(S_SNEQV) = col.sneqv;
    S(S_ZWT) = col.ZWT; S(S_WA) = col.WA; S(S_WT) = col.WT; S(S_TAUSS) = col.TAUSS;
    for (int k = 0; k < 7; ++k) { S(S_STC + k) = col.STC.v[k]; S(S_ZSNSO + k) = col.ZSNSO.v[k]; }
    for (int k = 0; k < 4; ++k) { S(S_SOILWAT + k) = col.soilwat.v[k]; S(S_SMC + k) = col.SMC.v[k]; }
    for (int k = 0; k < 3; ++k) { S(S_SNICE + k) = col.SNICE.v[k]; S(S_SNLIQ + k) = col.SNLIQ.v[k]; }
    S(S_QSFC) = col.QSFC; S(S_TAH) = col.TAH; S(S_CM) = col.CM; S(S_CH) = col.CH; S(S_WSLAKE) = col.WSLAKE;
    S(S_LFMASS) = col.LFMASS; S(S_RTMASS) = col.RTMASS; S(S_STMASS) = col.STMASS; S(S_WOOD) = col.WOOD;
    S(S_STBLCP) = col.STBLCP; S(S_FASTCP) = col.FASTCP; S(S_LAI) = col.LAI; S(S_SAI) = col.SAI;

    if (diag) {
      auto D = [&](int p) -> float& { return diag[(size_t)p * ncol + i]; };
      const float* d0 = &col.FSA;  // the 58 outputs are contiguous in ColumnIO in func:82-91 order
      for (int p = 0; p < 58; ++p) D(p) = d0[p];
      D(D_ERRWAT) = col.ERRWAT; D(D_ERRSW) = col.ERRSW; D(D_ERRENG) = col.ERRENG;
      D(D_ITER) = i2f((col.iter_vege & 0xff) | ((col.iter_stom_sun & 0xff) << 8) |
                      ((col.iter_stom_sha & 0xff) << 16) | ((col.niter_soil & 0xff) << 24));
      D(D_FLAGS) = i2f((int)anyflag);
      for (int q = 0; q < 6; ++q) D(D_ACCPRCP + q) = acc[q];
    }
  }
  return nflag;
}

// ---- leaf-routine wrappers for unit tests -------------------------------------------------------
void oracle_esat(float T, float* out4) { esat(T, out4[0], out4[1], out4[2], out4[3]); }

void oracle_rosr12(float* P, const float* A, const float* B, float* C, const float* D, float* DELTA,
                   int NTOP, int NSOIL_, int NSNOW_) {
  rosr12(P, A, B, C, D, DELTA, NTOP, NSOIL_, NSNOW_);
}

void oracle_combo(float* a4, const float* b4) {  // {DZ,WLIQ,WICE,T}
  combo(a4[0], a4[1], a4[2], a4[3], b4[0], b4[1], b4[2], b4[3]);
}

void oracle_tdfcnd(oracle_ctx* c, int sltyp, float SMC, float soilwat, float* DF) {
  tdfcnd(c->cx, sltyp, *DF, SMC, soilwat);
}

// out: FAB,FRE,FTD,FTI,FREV,FREG for band IB, then GDIR,BGAP,WGAP
void oracle_twostream(oracle_ctx* c, int lutyp, int IB, int IC, float COSZ, float VAI, float FWET, float T,
                      const float* albgrd2, const float* albgri2, const float* rho2, const float* tau2,
                      float fveg, float* out9) {
  A2 ALBGRD, ALBGRI, RHO, TAU, FAB, FRE, FTD, FTI, FREV, FREG;
  for (int b = 1; b <= 2; ++b) {
    ALBGRD(b) = albgrd2[b - 1]; ALBGRI(b) = albgri2[b - 1]; RHO(b) = rho2[b - 1]; TAU(b) = tau2[b - 1];
    FAB(b) = FRE(b) = FTD(b) = FTI(b) = FREV(b) = FREG(b) = 0.0f;
  }
  float GDIR = 0, BGAP = 0, WGAP = 0;
  twostream(c->cx, lutyp, IB, IC, COSZ, VAI, FWET, T, ALBGRD, ALBGRI, RHO, TAU, fveg, 1, FAB, FRE, FTD, FTI,
            GDIR, FREV, FREG, BGAP, WGAP);
  out9[0] = FAB(IB); out9[1] = FRE(IB); out9[2] = FTD(IB); out9[3] = FTI(IB); out9[4] = FREV(IB);
  out9[5] = FREG(IB); out9[6] = GDIR; out9[7] = BGAP; out9[8] = WGAP;
}

void oracle_stomata(oracle_ctx* c, int lutyp, float igs, float sfcprs, float sfctmp, float apar, float tv,
                    float ea, float ei, float o2, float co2, float foln, float btran, float rb, float* rs,
                    float* psn, int* niter) {
  stomata(c->cx, lutyp, igs, sfcprs, sfctmp, apar, tv, ea, ei, o2, co2, foln, btran, rb, *rs, *psn, *niter);
}

void oracle_frh2o(oracle_ctx* c, int sltyp, float TKELV, float SMC, float soilwat, float* FREE) {
  frh2o(c->cx, sltyp, *FREE, TKELV, SMC, soilwat);
}

int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
