"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the CPU parity oracle (oracle/src).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under noahmp_b200/ imports this module.  Parity pin: tests/golden/ref_*.npz, see oracle/src/common.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class RunParams(C.Structure):  # oracle/src/layout.h: struct RunParams
    _fields_ = [("julian0", C.c_double), ("dt", C.c_float), ("zsoil", C.c_float * 4), ("zref", C.c_float),
                ("co2_frac", C.c_float), ("o2_frac", C.c_float), ("foln", C.c_float), ("yearlen", C.c_int)]


LIBS = {"parity": "liboracle.so", "fast": "liboracle_fast.so", "strict": "liboracle_cr.so"}


def _variant(fast) -> str:
    return fast if isinstance(fast, str) else ("fast" if fast else "parity")


def build(fast=False) -> str:
    target = LIBS[_variant(fast)]
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)
    return os.path.join(HERE, target)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    def __init__(self, tables, opts, fast=False):
        path = os.path.join(HERE, LIBS[_variant(fast)])
        if not os.path.exists(path):
            path = build(fast)
        self.lib = lib = C.CDLL(path)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_run.restype = C.c_longlong
        lib.oracle_run.argtypes = [C.c_void_p, C.POINTER(RunParams), C.c_longlong, C.c_int, C.POINTER(C.c_float),
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                   C.POINTER(C.c_float), C.c_int]
        self.ctx = C.c_void_p(lib.oracle_create())
        p = tables.pack()
        self._keep = p
        opts_a = (C.c_int * 12)(*opts)
        assert lib.oracle_set_options(self.ctx, opts_a) == 0
        rc = lib.oracle_set_tables(self.ctx, _fp(p["gen"]), C.c_int(p["gen"].size), _fp(p["soil"]),
                                   C.c_int(p["soil"].shape[0]), _fp(p["color"]), C.c_int(p["color"].shape[0]),
                                   _fp(p["veg_f"]), _ip(p["veg_i"]), C.c_int(p["veg_f"].shape[0]), _ip(p["flags"]))
        assert rc == 0, rc

    def layout(self, what: int) -> int:
        return self.lib.oracle_layout(C.c_int(what))

    def num_threads(self) -> int:
        return self.lib.oracle_num_threads()

    def run(self, state, forcing, static_f, static_i, *, dt, zsoil, julian0, yearlen, zref=10.0,
            co2_frac=355e-6, o2_frac=0.209, foln=1.0, diag=None, nthreads=0):
        """Advance ``state`` [NSTATE][ncol] in place by forcing.shape[0] steps; returns #flagged columns."""
        assert state.dtype == np.float32 and state.flags.c_contiguous
        assert forcing.dtype == np.float32 and forcing.flags.c_contiguous
        nsteps, _, ncol = forcing.shape
        rp = RunParams(julian0, dt, (C.c_float * 4)(*[float(z) for z in zsoil]), zref, co2_frac, o2_frac, foln,
                       yearlen)
        dptr = _fp(diag) if diag is not None else C.POINTER(C.c_float)()
        return self.lib.oracle_run(self.ctx, C.byref(rp), ncol, nsteps, _fp(state), _fp(forcing), _fp(static_f),
                                   _ip(static_i), dptr, nthreads)

    def run_workload(self, w, forcing=None, diag=None, nthreads=0, state=None):
        from noahmp_b200 import synth
        if forcing is None:
            forcing = synth.forcing_all(w)
        st = w.state if state is None else state
        return self.run(st, forcing, w.static_f, w.static_i, dt=w.dt, zsoil=synth.ZSOIL, julian0=w.julian0,
                        yearlen=w.yearlen, zref=w.zref, co2_frac=w.co2_frac, o2_frac=w.o2_frac, foln=w.foln,
                        diag=diag, nthreads=nthreads)

    def __del__(self):
        try:
            self.lib.oracle_destroy(self.ctx)
        except Exception:
            pass
