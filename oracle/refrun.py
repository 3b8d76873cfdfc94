"""Runs the mechanically translated reference (oracle/f90_to_cpp.py -> oracle/_ref/libnoahmp_ref.so).

TEST INFRASTRUCTURE ONLY.  Used by tools/make_reference_golden.py (here, where /root/reference is mounted) to make the
fixtures tests/golden/ref_*.npz that pin the hand-written oracle; the product never imports this.

`RefModel.run` has the calling convention of `oracle.Oracle.run` (SoA planes in, SoA planes out) and supplies the same
driver duties per step (JULIAN, CO2AIR/O2AIR, PSFC = SFCPRS, FICEOLD from the snow layers at the start of the step), so
that the two can be compared plane by plane.  A column on which the reference would `stop` (its fatal budget checks)
is reported in the returned mask and left as it was before that step.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
LIB = os.path.join(REFDIR, "libnoahmp_ref.so")
LIB64 = os.path.join(REFDIR, "libnoahmp_ref64.so")     # the fp64 twin: same statements, every `real` a double
SIG = os.path.join(REFDIR, "noahmp_ref.json")


def build(ref_root: str = "/root/reference", force: bool = False, bounds: bool = False, keep_cpp: bool = False,
          fp64: bool = False) -> str:
    """Translate + compile.  Needs the reference sources; with only the prebuilt library present, returns it.
    The generated C++ is a derived copy of reference source: it is deleted once compiled (keep_cpp=True to inspect it);
    only the library and the signature table stay in oracle/_ref/ (git-ignored)."""
    src = os.path.join(ref_root, "core", "module_noahmp_func.f90")
    lib = LIB64 if fp64 else LIB
    if not os.path.exists(src):
        if os.path.exists(lib) and os.path.exists(SIG):
            return lib
        raise FileNotFoundError(f"{src} not found and no prebuilt {lib}")
    tool = os.path.join(HERE, "f90_to_cpp.py")
    stale = force or not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(tool), os.path.getmtime(src))
    if stale:
        from oracle import f90_to_cpp
        cpp = os.path.join(REFDIR, "noahmp_ref64.cpp" if fp64 else "noahmp_ref.cpp")
        f90_to_cpp.translate(ref_root, cpp, SIG, real_type="double" if fp64 else "float")
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w", "-o", lib, cpp]
        if bounds:
            cmd.insert(1, "-DF_BOUNDS")
        subprocess.run(cmd, check=True)
        if not keep_cpp:
            os.remove(cpp)
    return lib


_CT = {"real": C.c_float, "integer": C.c_int, "double": C.c_double, "logical": C.c_bool}
_NP = {"real": np.float32, "integer": np.int32, "double": np.float64, "logical": np.bool_}


class RefModel:
    def __init__(self, tables, opts, ref_root: str = "/root/reference", fp64: bool = False):
        """fp64=True loads the twin in which every `real` is a double (constants keep their fp32 values): only
        `run_native` may be used with it (planes stay fp32 at the boundary)."""
        self.fp64 = fp64
        self.lib = C.CDLL(build(ref_root, fp64=fp64))
        self.lib.ref_set.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        self.lib.ref_last_fatal.restype = C.c_char_p
        meta = json.load(open(SIG))
        self.sigs, self.modvars, self.iparams = meta["subs"], meta["module_vars"], meta["int_params"]
        self.set_options(opts)
        self.set_tables(tables)

    # ---- module variables ------------------------------------------------------------------------
    def _dims(self, name):
        out = []
        for d in self.modvars[name]["dims"]:
            d = d.strip()
            out.append(int(d) if d.isdigit() else self.iparams[d])
        return out

    def set_var(self, name, value):
        mv = self.modvars[name]
        dims = self._dims(name)
        arr = np.asarray(value, _NP[mv["type"]])
        if dims:
            full = np.full(dims, np.nan if mv["type"] != "integer" else -1, _NP[mv["type"]])
            sl = tuple(slice(0, n) for n in arr.shape)
            full[sl] = arr
            arr = np.asfortranarray(full)
            buf = arr.ravel(order="F")
        else:
            buf = arr.reshape(1)
        buf = np.ascontiguousarray(buf)
        rc = self.lib.ref_set(name.encode(), buf.ctypes.data, buf.size)
        if rc != 0:
            raise RuntimeError(f"ref_set({name}) -> {rc}")

    def set_options(self, opts):
        names = ["opt_veg", "opt_crs", "opt_btr", "opt_run", "opt_sfc", "opt_frz", "opt_inf", "opt_rad", "opt_alb", "opt_snf",
                 "opt_tbot", "opt_stc"]                                         # global:77-112
        for n, v in zip(names, opts):
            self.set_var(n, int(v))

    def set_tables(self, t):
        """Module variables as the reference's three *_readptable routines leave them (gen_param:63-88,
        soil_param:47-71, veg_param:89-160); the values come from noahmp_b200.tables.ParamTables."""
        g = t.gen
        for nm in ("CSOIL", "ZBOT", "CZIL", "DKREF", "KDTREF", "FRZK", "TIMEAN", "FSATMAX", "MLTFCT", "Z0SNO", "SSI", "SWEMAX",
                   "BETADS", "BETAIS", "EMSSOIL", "EMSLAKE"):
            self.set_var("kk_" + nm.lower(), g[nm])
        for nm in ("ALBICE", "ALBLAKE", "OMEGAS"):
            self.set_var("kk_" + nm.lower(), g[nm])
        self.set_var("lk_slope", g["SLOPE"])
        self.set_var("nslptyp", len(g["SLOPE"]))
        for nm in ("BEXP", "SMCMAX", "SMCREF", "SMCWLT", "PSISAT", "DKSAT", "DWSAT", "QUARTZ", "KDT", "FRZX"):
            self.set_var("lk_" + nm.lower(), t.soil[nm])
        self.set_var("nsltyp", t.nsoiltyp)
        self.set_var("nsoilcol", t.ncolor)
        for nm in ("ALBSAT", "ALBDRY"):
            self.set_var("lk_" + nm.lower(), np.stack([t.color[nm + "_VIS"], t.color[nm + "_NIR"]], axis=0))
        v = t.veg_f
        self.set_var("nlutyp", t.nvegtyp)
        for nm in ("RHOL", "RHOS", "TAUL", "TAUS"):
            self.set_var("lk_" + nm.lower(), np.stack([v[nm + "_VIS"], v[nm + "_NIR"]], axis=0))
        for nm in ("LAI", "SAI"):
            self.set_var(f"lk_{nm.lower()}12m", np.stack([v[f"{nm}{m:02d}"] for m in range(1, 13)], axis=0))
        skip = {"RHOL", "RHOS", "TAUL", "TAUS"}
        for nm in v:
            base = nm.split("_")[0]
            if base in skip or nm.startswith(("LAI", "SAI")) and nm[3:].isdigit():
                continue
            self.set_var("lk_" + nm.lower(), v[nm])
        for nm in ("NROOT", "C3C4"):
            self.set_var("lk_" + nm.lower(), t.veg_i[nm])
        for nm in ("ISURBAN", "ISWATER", "ISBARREN", "ISICE", "ISEGBLF"):
            self.set_var(nm.lower(), t.flags[nm])

    # ---- generic call ----------------------------------------------------------------------------
    def call(self, sub: str, values: dict):
        """Call a translated subroutine; `values` maps dummy-argument names to scalars / numpy arrays (arrays are
        updated in place, scalars are returned in the result dict)."""
        assert not self.fp64, "per-call interface is fp32 only"
        sig = self.sigs[sub]
        holders, ptrs = [], (C.c_void_p * len(sig))()
        for k, a in enumerate(sig):
            val = values.get(a["name"], 0)
            if a["rank"] > 0:
                arr = val
                assert isinstance(arr, np.ndarray) and arr.dtype == _NP[a["type"]] and arr.flags.c_contiguous, a["name"]
                holders.append(arr)
                ptrs[k] = arr.ctypes.data
            else:
                h = _CT[a["type"]](val)
                holders.append(h)
                ptrs[k] = C.cast(C.pointer(h), C.c_void_p)
        rc = getattr(self.lib, "ref_" + sub)(ptrs)
        out = {a["name"]: (h.value if a["rank"] == 0 else h) for a, h in zip(sig, holders)}
        return rc, out

    # ---- the column loop ---------------------------------------------------------------------------
    def run(self, state, forcing, static_f, static_i, dt, zsoil, julian0, yearlen, zref, co2_frac, o2_frac, foln, diag=None):
        """Advance every column through all steps; returns a bool mask of the columns the reference aborted on."""
        from noahmp_b200 import layout as L
        S, D = L.S, L.D
        nsteps, _, ncol = forcing.shape
        f32 = np.float32
        fatal = np.zeros(ncol, bool)
        zs = np.asarray(zsoil, f32).copy()
        scal_state = ["ALBOLD", "SNEQVO", "TAH", "EAH", "FWET", "CANLIQ", "CANICE", "TV", "TG", "QSFC", "QSNOW", "SNOWH", "SNEQV",
                      "ZWT", "WA", "WT", "WSLAKE", "LFMASS", "RTMASS", "STMASS", "WOOD", "STBLCP", "FASTCP", "LAI", "SAI", "CM",
                      "CH", "TAUSS"]
        out_names = [a["name"] for a in self.sigs["noahmp_sflx"] if a["intent"] == "out"]
        assert len(out_names) == 58
        for i in range(ncol):
            ty = int(static_i[i])
            col = {"iloc": 1, "jloc": 1, "nsoil": 4, "nsnow": 3, "dt": f32(dt), "dx": f32(0), "dz8w": f32(0), "iz0tlnd": 0,
                   "qc": f32(0), "pblh": f32(0), "foln": f32(foln), "yearlen": int(yearlen), "zlvl": f32(zref),
                   "lat": static_f[L.C["LAT"], i], "shdfac": static_f[L.C["SHDFAC"], i], "tbot": static_f[L.C["TBOT"], i],
                   "shdmax": static_f[L.C["SHDMAX"], i],
                   "lutyp": ty & 0xff, "sltyp": (ty >> 8) & 0xff, "slptyp": (ty >> 16) & 0xff, "isc": (ty >> 24) & 0xf,
                   "ist": ((ty >> 28) & 1) + 1, "ice": (ty >> 29) & 1, "zsoil": zs}
            for nm in scal_state:
                col[nm.lower()] = f32(state[S[nm], i])
            col["isnow"] = int(state[S["ISNOW"], i:i + 1].view(np.int32)[0])
            col["stc"] = np.ascontiguousarray(state[L.stc(-2):L.stc(4) + 1, i], f32)
            col["zsnso"] = np.ascontiguousarray(state[L.zsnso(-2):L.zsnso(4) + 1, i], f32)
            col["soilwat"] = np.ascontiguousarray(state[L.soilwat(1):L.soilwat(4) + 1, i], f32)
            col["smc"] = np.ascontiguousarray(state[L.smc(1):L.smc(4) + 1, i], f32)
            col["snice"] = np.ascontiguousarray(state[L.snice(-2):L.snice(0) + 1, i], f32)
            col["snliq"] = np.ascontiguousarray(state[L.snliq(-2):L.snliq(0) + 1, i], f32)
            outs = None
            for s in range(nsteps):
                f = forcing[s, :, i]
                col.update(sfctmp=f[L.F["SFCTMP"]], sfcprs=f[L.F["SFCPRS"]], psfc=f[L.F["SFCPRS"]], uu=f[L.F["UU"]], vv=f[L.F["VV"]],
                           q2=f[L.F["Q2"]], soldn=f[L.F["SOLDN"]], lwdn=f[L.F["LWDN"]], prcp=f[L.F["PRCP"]], cosz=f[L.F["COSZ"]])
                col["co2air"] = f32(co2_frac) * f32(col["sfcprs"])
                col["o2air"] = f32(o2_frac) * f32(col["sfcprs"])
                jd = np.float64(julian0) + np.float64(s) * np.float64(f32(dt)) / 86400.0
                col["julian"] = f32(jd - np.float64(yearlen) * np.floor(jd / np.float64(yearlen)))
                fice = np.zeros(3, f32)
                for k in (-2, -1, 0):
                    if k >= col["isnow"] + 1:
                        fice[k + 2] = col["snice"][k + 2] / (col["snice"][k + 2] + col["snliq"][k + 2])
                col["ficeold"] = fice
                keep = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in col.items()}
                rc, res = self.call("noahmp_sflx", col)
                if rc != 0:
                    fatal[i] = True
                    col = keep
                    break
                for a in self.sigs["noahmp_sflx"]:
                    if a["rank"] == 0 and a["intent"] in ("inout", "out"):
                        col[a["name"]] = res[a["name"]]
                outs = res
            for nm in scal_state:
                state[S[nm], i] = col[nm.lower()]
            state[S["ISNOW"], i:i + 1].view(np.int32)[0] = col["isnow"]
            state[L.stc(-2):L.stc(4) + 1, i] = col["stc"]
            state[L.zsnso(-2):L.zsnso(4) + 1, i] = col["zsnso"]
            state[L.soilwat(1):L.soilwat(4) + 1, i] = col["soilwat"]
            state[L.smc(1):L.smc(4) + 1, i] = col["smc"]
            state[L.snice(-2):L.snice(0) + 1, i] = col["snice"]
            state[L.snliq(-2):L.snliq(0) + 1, i] = col["snliq"]
            if diag is not None and outs is not None:
                for p, nm in enumerate(out_names):
                    diag[p, i] = outs[nm]
        return fatal

    def run_native(self, state, forcing, static_f, static_i, dt, zsoil, julian0, yearlen, zref, co2_frac, o2_frac, foln,
                   diag=None, nthreads=0):
        """Same contract as `run`, but the column loop is the generated C++ one (OpenMP): for big comparisons and timing."""
        from noahmp_b200 import layout as L
        from oracle.f90_to_cpp import NATIVE_ARRAYS, NATIVE_FORCING, NATIVE_SCALARS, NATIVE_STATIC

        class RP(C.Structure):
            _fields_ = [("julian0", C.c_double), ("dt", C.c_float), ("zsoil", C.c_float * 4), ("zref", C.c_float),
                        ("co2_frac", C.c_float), ("o2_frac", C.c_float), ("foln", C.c_float), ("yearlen", C.c_int)]
        nsteps, _, ncol = forcing.shape
        assert state.flags.c_contiguous and forcing.flags.c_contiguous and state.dtype == np.float32
        rp = RP(float(julian0), float(dt), (C.c_float * 4)(*[float(z) for z in zsoil]), float(zref), float(co2_frac), float(o2_frac),
                float(foln), int(yearlen))
        ia = lambda xs: np.asarray(xs, np.int32)                                 # noqa: E731
        sp = ia([L.S[n.upper()] for n in NATIVE_SCALARS])
        base = {"stc": L.stc(-2), "zsnso": L.zsnso(-2), "soilwat": L.soilwat(1), "smc": L.smc(1), "snice": L.snice(-2), "snliq": L.snliq(-2)}
        ap = ia([base[n] for n, _ in NATIVE_ARRAYS])
        fp = ia([L.F[n.upper()] for n in NATIVE_FORCING])
        cp = ia([L.C[n.upper()] for n in NATIVE_STATIC])
        fatal = np.zeros(ncol, np.uint8)
        sf = np.ascontiguousarray(static_f, np.float32)
        si = np.ascontiguousarray(static_i, np.int32)
        if nthreads:
            os.environ["OMP_NUM_THREADS"] = str(nthreads)
        fn = self.lib.ref_run_columns
        fn.restype = C.c_longlong
        fn.argtypes = [C.c_longlong, C.c_int] + [C.c_void_p] * 6 + [C.POINTER(RP)] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 2
        fn(ncol, nsteps, state.ctypes.data, forcing.ctypes.data, sf.ctypes.data, si.ctypes.data,
           diag.ctypes.data if diag is not None else None, fatal.ctypes.data, C.byref(rp), sp.ctypes.data, ap.ctypes.data,
           int(L.S["ISNOW"]), fp.ctypes.data, cp.ctypes.data)
        return fatal.astype(bool)

    def last_fatal(self) -> str:
        return self.lib.ref_last_fatal().decode()
