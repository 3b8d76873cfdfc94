"""ctypes binding of the C ABI (``include/noahmp_b200.h``).  There is no CPU fallback: if the CUDA
library is missing or does not load, importing users get a hard error."""
from __future__ import annotations

import ctypes as C
import os

from .build import lib_path


class RunParams(C.Structure):  # struct noahmp_run_params
    _fields_ = [("julian0", C.c_double), ("dt", C.c_float), ("zsoil", C.c_float * 4), ("zref", C.c_float),
                ("co2_frac", C.c_float), ("o2_frac", C.c_float), ("foln", C.c_float), ("yearlen", C.c_int)]


class StepForcing(C.Structure):  # struct noahmp_step_forcing
    _fields_ = [("rec0", C.c_int), ("rec1", C.c_int), ("weight", C.c_float), ("sd", C.c_float), ("cd", C.c_float),
                ("ch0", C.c_float), ("sh0", C.c_float)]


# every symbol declared in include/noahmp_b200.h
SYMBOLS = ["noahmp_init", "noahmp_finalize", "noahmp_set_stream", "noahmp_last_error", "noahmp_set_options",
           "noahmp_set_tables", "noahmp_run", "noahmp_run_host", "noahmp_advance_host", "noahmp_run_records",
           "noahmp_advance_records_host", "noahmp_synth_forcing", "noahmp_synth_forcing_ids", "noahmp_prepare_forcing", "noahmp_layout",
           "noahmp_state_plane_name", "noahmp_forcing_plane_name", "noahmp_diag_plane_name", "noahmp_budget_name",
           "noahmp_set_record_packing", "noahmp_pack_plane", "noahmp_h2d_bytes", "noahmp_pack_seconds",
           "noahmp_launch_count", "noahmp_kernel_info", "noahmp_debug_bounds", "noahmp_debug_bounds_selftest"]

_cache: dict = {}


class NoahMPLibraryError(RuntimeError):
    pass


def load(fast=False) -> C.CDLL:
    """Load the in-tree CUDA library.  Raises NoahMPLibraryError (never falls back) when absent."""
    key = fast if isinstance(fast, str) else bool(fast)
    if key in _cache:
        return _cache[key]
    path = lib_path(fast)
    if not os.path.exists(path):
        raise NoahMPLibraryError(
            f"{path} not found: build it with `python -m noahmp_b200.build --all` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    try:
        lib = C.CDLL(path)
    except OSError as exc:  # e.g. no libcudart / no driver
        raise NoahMPLibraryError(f"cannot load {path}: {exc}") from exc
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.noahmp_init.argtypes = [C.POINTER(vp), C.c_int]
    lib.noahmp_finalize.argtypes = [vp]
    lib.noahmp_set_stream.argtypes = [vp, vp]
    lib.noahmp_last_error.argtypes = [vp]
    lib.noahmp_last_error.restype = C.c_char_p
    lib.noahmp_set_options.argtypes = [vp, ip]
    lib.noahmp_set_tables.argtypes = [vp, fp, C.c_int, fp, C.c_int, fp, C.c_int, fp, ip, C.c_int, ip]
    lib.noahmp_run.argtypes = [vp, C.POINTER(RunParams), C.c_int64, C.c_int, vp, vp, C.c_int64, vp, vp, vp, vp]
    lib.noahmp_run_host.argtypes = [vp, C.POINTER(RunParams), C.c_int64, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.noahmp_run_host.restype = C.c_int64
    lib.noahmp_advance_host.argtypes = [vp, C.POINTER(RunParams), C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    lib.noahmp_advance_host.restype = C.c_int64
    sfp = C.POINTER(StepForcing)
    lib.noahmp_run_records.argtypes = [vp, C.POINTER(RunParams), C.c_int64, C.c_int, vp, vp, C.c_int64, sfp, vp, vp, vp, vp, vp]
    lib.noahmp_advance_records_host.argtypes = [vp, C.POINTER(RunParams), C.c_int64, C.c_int, vp, vp, C.c_int, sfp, vp, vp, vp, vp, C.c_int]
    lib.noahmp_advance_records_host.restype = C.c_int64
    lib.noahmp_synth_forcing.argtypes = [vp, C.c_int64, C.c_int64, C.c_uint32, C.c_int, fp, fp, vp, vp]
    lib.noahmp_synth_forcing_ids.argtypes = [vp, C.c_int64, vp, C.c_uint32, C.c_int, fp, fp, vp, vp]
    lib.noahmp_prepare_forcing.argtypes = [vp, C.c_int64, vp, vp, C.c_float, vp, fp, vp]
    lib.noahmp_layout.argtypes = [C.c_int]
    for f in ("noahmp_state_plane_name", "noahmp_forcing_plane_name", "noahmp_diag_plane_name", "noahmp_budget_name"):
        getattr(lib, f).argtypes = [C.c_int]
        getattr(lib, f).restype = C.c_char_p
    lib.noahmp_set_record_packing.argtypes = [vp, C.c_uint, C.c_int]
    lib.noahmp_pack_plane.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int]
    lib.noahmp_pack_plane.restype = C.c_int64
    lib.noahmp_h2d_bytes.argtypes = [vp]
    lib.noahmp_h2d_bytes.restype = C.c_int64
    lib.noahmp_pack_seconds.argtypes = [vp]
    lib.noahmp_pack_seconds.restype = C.c_double
    lib.noahmp_launch_count.argtypes = [vp]
    lib.noahmp_launch_count.restype = C.c_int64
    lib.noahmp_kernel_info.argtypes = [vp, ip]
    lib.noahmp_debug_bounds.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.c_int]
    lib.noahmp_debug_bounds_selftest.argtypes = [vp, C.c_int]
    _cache[key] = lib
    return lib
