"""Host-side engine: the Python mirror of the reference's (empty) ``noahmp_engine`` module.

``core/module_noahmp_engine.f90:5-10`` declares ``noahmp_init`` and ``noahmp_run`` without bodies; this
module keeps those two names as the operator surface and implements them over the CUDA C ABI.  PyTorch
is used only to own device buffers and streams -- tensors are handed to the library as raw pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import layout as L
from ._lib import NoahMPLibraryError, RunParams, StepForcing, load
from .tables import ParamTables

OPTION_ORDER = ["opt_veg", "opt_crs", "opt_btr", "opt_run", "opt_sfc", "opt_frz", "opt_inf", "opt_rad", "opt_alb",
                "opt_snf", "opt_tbot", "opt_stc"]  # core/module_noahmp_global.f90:77-80


class NoahMPError(RuntimeError):
    pass


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dev_ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda or not t.is_contiguous():
        raise NoahMPError("device entry points need contiguous CUDA tensors")
    return C.c_void_p(t.data_ptr())


def _host_ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    if hasattr(a, "data_ptr"):  # pinned torch tensor
        if a.is_cuda or not a.is_contiguous():
            raise NoahMPError("host entry point needs contiguous CPU buffers")
        return C.c_void_p(a.data_ptr())
    if not a.flags.c_contiguous:
        raise NoahMPError("host entry point needs C-contiguous arrays")
    return C.c_void_p(a.ctypes.data)


class NoahMP:
    """One context per device (calls on a context are stream-ordered, not thread-safe)."""

    def __init__(self, tables: ParamTables | None = None, opts=None, device: int = 0, fast: bool = False):
        self.lib = load(fast)
        self.fast = fast
        self.device = device
        ctx = C.c_void_p()
        rc = self.lib.noahmp_init(C.byref(ctx), device)
        if rc != 0:
            raise NoahMPError(f"noahmp_init failed on device {device} (code {rc}); a B200 GPU is required")
        self.ctx = ctx
        self.set_options(opts or [1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1])
        self.set_tables(tables or ParamTables.default())

    # -- configuration -----------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc < 0:
            raise NoahMPError(f"{what}: {self.lib.noahmp_last_error(self.ctx).decode()} (code {rc})")
        return rc

    def set_options(self, opts) -> None:
        """``noahmp_set_options`` (global:77-112): 12 ints in the reference's order, or a dict of opt_*."""
        if isinstance(opts, dict):
            opts = [opts[k] for k in OPTION_ORDER]
        self.opts = [int(o) for o in opts]
        self._check(self.lib.noahmp_set_options(self.ctx, (C.c_int * 12)(*self.opts)), "noahmp_set_options")

    def set_tables(self, tables: ParamTables) -> None:
        p = tables.pack()
        self._tables = p
        self._check(self.lib.noahmp_set_tables(
            self.ctx, _fp(p["gen"]), p["gen"].size, _fp(p["soil"]), p["soil"].shape[0], _fp(p["color"]),
            p["color"].shape[0], _fp(p["veg_f"]), _ip(p["veg_i"]), p["veg_f"].shape[0], _ip(p["flags"])),
            "noahmp_set_tables")

    def set_stream(self, stream) -> None:
        """``stream``: a torch.cuda.Stream (or None for the default stream)."""
        self._check(self.lib.noahmp_set_stream(self.ctx, C.c_void_p(stream.cuda_stream if stream is not None else 0)),
                    "noahmp_set_stream")

    @staticmethod
    def run_params(dt, zsoil, julian0, yearlen, zref=10.0, co2_frac=355e-6, o2_frac=0.209, foln=1.0) -> RunParams:
        return RunParams(float(julian0), float(dt), (C.c_float * 4)(*[float(z) for z in zsoil]), float(zref),
                         float(co2_frac), float(o2_frac), float(foln), int(yearlen))

    # -- the hot path ----------------------------------------------------------------------------------
    def run(self, rp: RunParams, state, forcing, static_f, static_i, diag=None, budget=None) -> None:
        """``noahmp_run`` on device tensors: state [NSTATE,ncol] f32 (in place), forcing [nsteps,9,ncol] f32,
        static_f [4,ncol] f32, static_i [ncol] i32, diag [NDIAG,ncol] f32 or None, budget [16] f64 or None.
        Asynchronous on the context's stream."""
        nsteps, nf, ncol = forcing.shape
        assert nf == L.NFORCING and state.shape == (L.NSTATE, ncol)
        self._check(self.lib.noahmp_run(self.ctx, C.byref(rp), ncol, nsteps, _dev_ptr(state), _dev_ptr(forcing),
                                        L.NFORCING * ncol, _dev_ptr(static_f), _dev_ptr(static_i), _dev_ptr(diag),
                                        _dev_ptr(budget)), "noahmp_run")

    def run_host(self, rp: RunParams, state, forcing, static_f, static_i, diag=None, budget=None) -> int:
        """``noahmp_run_host``: same call with host buffers (numpy arrays or pinned CPU tensors); copies
        in, runs, copies back, synchronises.  Returns the number of flagged columns."""
        nsteps, nf, ncol = forcing.shape
        assert nf == L.NFORCING and tuple(state.shape) == (L.NSTATE, ncol)
        return self._check(self.lib.noahmp_run_host(self.ctx, C.byref(rp), ncol, nsteps, _host_ptr(state),
                                                    _host_ptr(forcing), _host_ptr(static_f), _host_ptr(static_i),
                                                    _host_ptr(diag), _host_ptr(budget)), "noahmp_run_host")

    def advance_host(self, rp: RunParams, state, forcing_host, static_f, static_i, budget=None, steps_per_launch: int = 1) -> int:
        """``noahmp_advance_host``: device-resident ``state``/``static_*`` (CUDA tensors), forcing streamed from a
        (pinned) host buffer [nsteps,9,ncol] through a double buffer, H2D overlapped with compute.  ``budget``: host
        fp64 array/tensor [16] or None.  Returns the flagged column-launch count."""
        nsteps, nf, ncol = forcing_host.shape
        assert nf == L.NFORCING and tuple(state.shape) == (L.NSTATE, ncol)
        return self._check(self.lib.noahmp_advance_host(self.ctx, C.byref(rp), ncol, nsteps, _dev_ptr(state),
                                                        _host_ptr(forcing_host), _dev_ptr(static_f), _dev_ptr(static_i),
                                                        _host_ptr(budget), int(steps_per_launch)), "noahmp_advance_host")

    @staticmethod
    def step_table(rec0, rec1, weight, scal4):
        """``noahmp_step_forcing[nsteps]`` from per-step sequences: record indices, weights in [0,1] and the four solar
        scalars (sin / cos of the declination, cos / sin of the hour angle at longitude 0; ``forcing.solar_scalars``)."""
        n = len(rec0)
        tab = (StepForcing * n)()
        for s in range(n):
            sc = scal4[s]
            tab[s] = StepForcing(int(rec0[s]), int(rec1[s]), float(weight[s]), float(sc[0]), float(sc[1]), float(sc[2]), float(sc[3]))
        return tab

    def run_records(self, rp: RunParams, state, records, steps, geo, static_f, static_i, diag=None, budget=None) -> None:
        """``noahmp_run_records``: like :meth:`run`, the forcing given as device records [nrec,8,ncol] + a step table
        (:meth:`step_table`) + geo planes [4,ncol]; interpolation and COSZ happen in the step kernel's load."""
        nrec, nv, ncol = records.shape
        assert nv == 8 and state.shape == (L.NSTATE, ncol) and tuple(geo.shape) == (4, ncol)
        assert all(0 <= s.rec0 < nrec and 0 <= s.rec1 < nrec for s in steps)
        self._check(self.lib.noahmp_run_records(self.ctx, C.byref(rp), ncol, len(steps), _dev_ptr(state), _dev_ptr(records), 8 * ncol,
                                                steps, _dev_ptr(geo), _dev_ptr(static_f), _dev_ptr(static_i), _dev_ptr(diag),
                                                _dev_ptr(budget)), "noahmp_run_records")

    def advance_records_host(self, rp: RunParams, state, records_host, steps, geo, static_f, static_i, budget=None,
                             steps_per_launch: int = 1) -> int:
        """``noahmp_advance_records_host``: device-resident state / static / geo planes, records [nrec,8,ncol] streamed from
        (pinned) host memory, each copied once.  Returns the flagged column-launch count."""
        nrec, nv, ncol = records_host.shape
        assert nv == 8 and tuple(state.shape) == (L.NSTATE, ncol) and tuple(geo.shape) == (4, ncol)
        return self._check(self.lib.noahmp_advance_records_host(self.ctx, C.byref(rp), ncol, len(steps), _dev_ptr(state),
                                                                _host_ptr(records_host), nrec, steps, _dev_ptr(geo),
                                                                _dev_ptr(static_f), _dev_ptr(static_i), _host_ptr(budget),
                                                                int(steps_per_launch)), "noahmp_advance_records_host")

    def synth_forcing(self, ncol, col0, seed, step, coef17: np.ndarray, scalars4: np.ndarray, sincos, out) -> None:
        self._check(self.lib.noahmp_synth_forcing(self.ctx, ncol, col0, seed, step, _fp(coef17), _fp(scalars4),
                                                  _dev_ptr(sincos), _dev_ptr(out)), "noahmp_synth_forcing")

    def synth_forcing_ids(self, ids, seed, step, coef17: np.ndarray, scalars4: np.ndarray, sincos, out) -> None:
        """The generator for re-ordered columns: ``ids`` int64 device tensor [ncol] = generator column ids."""
        assert ids.dtype.is_floating_point is False and ids.element_size() == 8
        self._check(self.lib.noahmp_synth_forcing_ids(self.ctx, ids.shape[0], _dev_ptr(ids), seed, step, _fp(coef17), _fp(scalars4),
                                                      _dev_ptr(sincos), _dev_ptr(out)), "noahmp_synth_forcing_ids")

    def prepare_forcing(self, rec0, rec1, weight: float, geo, scal4: np.ndarray, out) -> None:
        """``noahmp_prepare_forcing``: out[9,ncol] = time interpolation of two [8,ncol] records + COSZ (device tensors)."""
        ncol = out.shape[1]
        assert tuple(rec0.shape) == (8, ncol) and tuple(rec1.shape) == (8, ncol) and tuple(geo.shape) == (4, ncol)
        self._check(self.lib.noahmp_prepare_forcing(self.ctx, ncol, _dev_ptr(rec0), _dev_ptr(rec1), float(weight), _dev_ptr(geo),
                                                    _fp(np.ascontiguousarray(scal4, np.float32)), _dev_ptr(out)),
                    "noahmp_prepare_forcing")

    def set_record_packing(self, plane_mask: int = 0xA0, host_threads: int = 0) -> None:
        """``noahmp_set_record_packing``: which planes of a record :meth:`advance_records_host` zero-suppresses on the host
        before the copy (default SOLDN | PRCP; 0 = off) and with how many host threads (0 = default)."""
        self._check(self.lib.noahmp_set_record_packing(self.ctx, int(plane_mask), int(host_threads)), "noahmp_set_record_packing")

    # -- introspection -----------------------------------------------------------------------------------
    def h2d_bytes(self) -> int:
        """Bytes the ``advance_*`` entry points of this context have copied host -> device so far."""
        return int(self.lib.noahmp_h2d_bytes(self.ctx))

    def pack_seconds(self) -> float:
        """Wall time ``advance_records_host`` has spent zero-suppressing record planes on the host so far."""
        return float(self.lib.noahmp_pack_seconds(self.ctx))

    def launch_count(self) -> int:
        return int(self.lib.noahmp_launch_count(self.ctx))

    def kernel_info(self) -> dict:
        info = (C.c_int * 4)()
        self._check(self.lib.noahmp_kernel_info(self.ctx, info), "noahmp_kernel_info")
        return {"registers": info[0], "local_bytes": info[1], "static_smem": info[2], "max_threads": info[3]}

    def debug_bounds(self, reset: bool = False) -> dict:
        """Counts of the device-side index checks of a ``-DNMP_BOUNDS`` library (``fast="bounds"``); ``checked`` is False for
        every other build."""
        hits, sites = C.c_ulonglong(0), C.c_uint(0)
        rc = self._check(self.lib.noahmp_debug_bounds(self.ctx, C.byref(hits), C.byref(sites), int(reset)), "noahmp_debug_bounds")
        return {"checked": rc == 1, "hits": int(hits.value), "sites": int(sites.value)}

    def debug_bounds_selftest(self, idx: int) -> bool:
        return self._check(self.lib.noahmp_debug_bounds_selftest(self.ctx, int(idx)), "noahmp_debug_bounds_selftest") == 1

    def close(self) -> None:
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.noahmp_finalize(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the reference's two entry-point names (engine:5-10) ---------------------------------------------
_default: NoahMP | None = None


def pack_plane(plane: np.ndarray, host_threads: int = 1, fast="fast"):
    """``noahmp_pack_plane`` (host only, no device): the zero-suppressed form of one record plane that
    ``noahmp_advance_records_host`` sends over the host link.  Returns (occupancy uint32[G], before uint32[G],
    values float32[nnz]) with G = ceil(ncol / 32)."""
    plane = np.ascontiguousarray(plane, np.float32)
    n = plane.shape[0]
    g = (n + 31) // 32
    occ, bef, val = np.zeros(g, np.uint32), np.zeros(g, np.uint32), np.zeros(max(n, 1), np.float32)
    nnz = load(fast).noahmp_pack_plane(plane.ctypes.data, n, occ.ctypes.data, bef.ctypes.data, val.ctypes.data, int(host_threads))
    if nnz < 0:
        raise RuntimeError("noahmp_pack_plane failed")
    return occ, bef, val[:nnz].copy()


def noahmp_init(tables: ParamTables | None = None, opts=None, device: int = 0, fast: bool = False) -> NoahMP:
    """Replaces the empty ``subroutine noahmp_init``: options + the three table reads + device context."""
    global _default
    _default = NoahMP(tables, opts, device, fast)
    return _default


def noahmp_run(rp, state, forcing, static_f, static_i, diag=None, budget=None, engine: NoahMP | None = None) -> None:
    """Replaces the empty ``subroutine noahmp_run``: advance every column by ``forcing.shape[0]`` steps."""
    eng = engine or _default
    if eng is None:
        raise NoahMPError("noahmp_init has not been called")
    eng.run(rp, state, forcing, static_f, static_i, diag, budget)


__all__ = ["NoahMP", "NoahMPError", "NoahMPLibraryError", "noahmp_init", "noahmp_run", "OPTION_ORDER"]
