"""Structure-of-arrays plane names of the C ABI (``include/noahmp_b200.h``).

Every per-column quantity of the reference's 131-argument ``noahmp_sflx`` (``core/module_noahmp_func.f90:66-91``,
classified in SURVEY App. A) lives in one fp32 plane ``[plane][ncol]``; integers are stored as their
int32 bit pattern in a plane of the same buffer.
"""
from __future__ import annotations

# ---- state (in/out) ------------------------------------------------------------------------------
STATE_SCALARS = ["ALBOLD", "SNEQVO", "EAH", "FWET", "CANLIQ", "CANICE", "TV", "TG", "QSNOW", "ISNOW",
                 "SNOWH", "SNEQV", "ZWT", "WA", "WT", "TAUSS"]
STATE_CORE = (STATE_SCALARS
              + [f"STC{k:+d}" for k in range(-2, 5)]
              + [f"SOILWAT{k}" for k in range(1, 5)]
              + [f"SMC{k}" for k in range(1, 5)]
              + [f"ZSNSO{k:+d}" for k in range(-2, 5)]
              + [f"SNICE{k:+d}" for k in range(-2, 1)]
              + [f"SNLIQ{k:+d}" for k in range(-2, 1)])
STATE_EXT = ["QSFC", "TAH", "CM", "CH", "WSLAKE", "LFMASS", "RTMASS", "STMASS", "WOOD", "STBLCP", "FASTCP",
             "LAI", "SAI"]
STATE = STATE_CORE + STATE_EXT
NSTATE_CORE = len(STATE_CORE)  # 44
NSTATE = len(STATE)            # 57
S = {n: i for i, n in enumerate(STATE)}
S_STC, S_SOILWAT, S_SMC, S_ZSNSO, S_SNICE, S_SNLIQ = 16, 23, 27, 31, 38, 41
INT_STATE_PLANES = [S["ISNOW"]]


def stc(j: int) -> int:
    """Plane of STC(j), Fortran layer index j in -2..4 (snow -2..0, soil 1..4)."""
    return S_STC + j + 2


def zsnso(j: int) -> int:
    return S_ZSNSO + j + 2


def snice(j: int) -> int:
    return S_SNICE + j + 2


def snliq(j: int) -> int:
    return S_SNLIQ + j + 2


def smc(k: int) -> int:
    """Plane of SMC(k), soil layer k in 1..4."""
    return S_SMC + k - 1


def soilwat(k: int) -> int:
    return S_SOILWAT + k - 1

# ---- forcing (read every step) -------------------------------------------------------------------
FORCING = ["SFCTMP", "SFCPRS", "UU", "VV", "Q2", "SOLDN", "LWDN", "PRCP", "COSZ"]
NFORCING = len(FORCING)
F = {n: i for i, n in enumerate(FORCING)}

# ---- static per column ---------------------------------------------------------------------------
STATIC_F = ["LAT", "SHDFAC", "TBOT", "SHDMAX"]
NSTATIC_F = len(STATIC_F)
C = {n: i for i, n in enumerate(STATIC_F)}

# ---- diagnostics: the 58 outputs in the order of func:82-91, then residuals / counters / flags -----
DIAG = ["FSA", "FSR", "FIRA", "FSH", "SSOIL", "FCEV", "FGEV", "FCTR", "ECAN", "ETRAN", "EDIR", "TRAD",
        "TGB", "TGV", "T2MV", "T2MB", "Q2V", "Q2B", "RUNSRF", "RUNSUB", "APAR", "PSN", "SAV", "SAG", "FSNO",
        "NEE", "GPP", "NPP", "FVEG", "ALBEDO", "QSNBOT", "PONDING", "PONDING1", "PONDING2", "RSSUN",
        "RSSHA", "BGAP", "WGAP", "CHV", "CHB", "EMISSI", "SHG", "SHC", "SHB", "EVG", "EVB", "GHV", "GHB",
        "IRG", "IRC", "IRB", "TR", "EVC", "CHLEAF", "CHUC", "CHV2", "CHB2", "FPICE",
        "ERRWAT", "ERRSW", "ERRENG", "ITER", "FLAGS",
        # accumulated since last zeroed by the caller (mm), HRLDAS names: read-modify-write planes
        "ACCPRCP", "ACCECAN", "ACCETRAN", "ACCEDIR", "SFCRNOFF", "UGDRNOFF"]
NDIAG = len(DIAG)  # 69
ACC_DIAG_PLANES = list(range(63, 69))
D = {n: i for i, n in enumerate(DIAG)}
INT_DIAG_PLANES = [D["ITER"], D["FLAGS"]]

# ---- error flag bits (one per fatal site of the reference) ----------------------
FLAG_BITS = {"ERRSW": 1 << 0, "ERRENG": 1 << 1, "FIRE": 1 << 2, "HCAN": 1 << 3, "ZLVL": 1 << 4,
             "OPTVEG": 1 << 5, "NAN": 1 << 6, "BADTYPE": 1 << 7}

# ---- budget sums reduced on device (and all-reduced over ranks), fp64 ------------------------------
BUDGET = ["NCELL", "NFLAGGED", "ERRWAT_SUM", "ERRWAT_ABS", "ERRWAT_MAXABS", "FSA", "FSR", "FIRA", "FSH", "LH",
          "SSOIL", "PRCP_DT", "ET_DT", "RUNOFF_DT", "SNEQV", "SOILWATER"]
NBUDGET = len(BUDGET)  # 16
B = {n: i for i, n in enumerate(BUDGET)}


def pack_types(lutyp, sltyp, slptyp, isc, ist, ice):
    """lutyp | sltyp<<8 | slptyp<<16 | ISC<<24 | (IST-1)<<28 | ICE<<29  (works on ints or numpy arrays)."""
    return ((lutyp & 0xff) | ((sltyp & 0xff) << 8) | ((slptyp & 0xff) << 16) | ((isc & 0xf) << 24)
            | (((ist - 1) & 1) << 28) | ((ice & 1) << 29))
