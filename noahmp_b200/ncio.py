"""NetCDF input and output for the offline driver.

The reference's namelist names netCDF files (run/case.nml: ``geo_em.d01.nc``, ``init.nc``, ``restart.nc``; forcing in
``ldasin``, output in ``ldasout``) but contains no reader or writer.  This image has no netCDF4/HDF5 library; NetCDF-3
(classic and 64-bit offset) is read and written through ``scipy.io.netcdf_file``.  A NetCDF-4 file is refused with a message
that says how to convert it (``nccopy -k classic``).

Conventions accepted (HRLDAS / WRF names first, this repo's own names second):

* initialisation / static file: ``XLAT XLONG IVGTYP ISLTYP TSK TSLB|STC SMOIS|SMC SH2O SNOWH SNOW|SNEQV SHDFAC|VEGFRA TMN
  CANWAT ZWT WA SLOPETYP|SLPTYP ISC SHDMAX``; any of ``[Time,] [soil,] south_north, west_east`` or already flat
  ``[soil,] ncol``.  Grids are flattened row-major to columns; the grid shape is kept and used for the files written.
* forcing record ``<YYYYMMDDHH>.LDASIN_DOMAIN1`` / ``<YYYYMMDDHHMM>.nc``: ``T2D PSFC U2D V2D Q2D SWDOWN LWDOWN RAINRATE``
  (or ``SFCTMP SFCPRS UU VV Q2 SOLDN LWDN PRCP``), one time level.
* output ``<YYYYMMDDHHMM>.LDASOUT_DOMAIN1`` (one variable per diagnostic plane) and restart
  ``RESTART.<YYYYMMDDHHMM>_DOMAIN1`` (one variable per state plane; ``ISNOW`` as int32), same grid.
"""
from __future__ import annotations

import datetime as dt
import os
import sys

import numpy as np

from . import layout as L

f32 = np.float32

INIT_ALIASES = {
    "lat": ("XLAT", "lat", "LAT"), "lon": ("XLONG", "lon", "LON"), "lutyp": ("IVGTYP", "lutyp", "LU_INDEX"),
    "sltyp": ("ISLTYP", "sltyp"), "slptyp": ("SLOPETYP", "SLPTYP", "slptyp"), "isc": ("ISC", "isc", "SOILCOLOR"),
    "shdfac": ("SHDFAC", "VEGFRA", "shdfac"), "shdmax": ("SHDMAX", "GVFMAX", "shdmax"), "tbot": ("TMN", "TBOT", "tbot"),
    "tsk": ("TSK", "tsk"), "stc": ("TSLB", "STC", "SOIL_T", "stc"), "smc": ("SMOIS", "SMC", "SOIL_M", "smc"),
    "sh2o": ("SH2O", "SOIL_W", "sh2o"), "snowh": ("SNOWH", "snowh"), "sneqv": ("SNOW", "SNEQV", "WEASD", "sneqv"),
    "canliq": ("CANWAT", "CANLIQ", "canliq"), "zwt": ("ZWT", "zwt"), "wa": ("WA", "wa"), "eah": ("EAH", "eah"),
}
FORCING_ALIASES = [("T2D", "SFCTMP"), ("PSFC", "SFCPRS"), ("U2D", "UU"), ("V2D", "VV"), ("Q2D", "Q2"), ("SWDOWN", "SOLDN"),
                   ("LWDOWN", "LWDN"), ("RAINRATE", "PRCP")]
_SOIL_KEYS = ("stc", "smc", "sh2o")
_PERCENT_KEYS = ("shdfac", "shdmax")    # VEGFRA / GVFMAX are percentages in WRF files


def is_netcdf(path: str) -> bool:
    """True for a NetCDF-3 file; raises for NetCDF-4/HDF5 (not readable here)."""
    try:
        with open(path, "rb") as fh:
            magic = fh.read(8)
    except OSError:
        return False
    if magic[:3] == b"CDF" and magic[3:4] in (b"\x01", b"\x02"):
        return True
    if magic == b"\x89HDF\r\n\x1a\n":
        raise ValueError(f"ERR: {path} is NetCDF-4/HDF5; this build reads NetCDF-3 only - convert it with "
                         f"`nccopy -k classic {os.path.basename(path)} out.nc`")
    return False


def _open(path, mode="r"):
    from scipy.io import netcdf_file
    return netcdf_file(path, mode, mmap=False, version=2)


def _squeeze_time(v):
    """Variable data without a leading Time dimension of length 1 (copy: the file is closed afterwards)."""
    a = np.array(v.data)
    dims = list(v.dimensions)
    while a.ndim > 1 and dims and dims[0].lower() in ("time", "times") and a.shape[0] == 1:
        a, dims = a[0], dims[1:]
    return a, dims


def _flatten(a, dims, soil: bool):
    """[..grid..] -> [ncol] or, for soil variables, [nsoil][ncol] whatever the position of the soil dimension."""
    if soil:
        k = next((i for i, d in enumerate(dims) if "soil" in d.lower() or d.lower() in ("layer", "layers", "nsoil")), 0)
        a = np.moveaxis(a, k, 0)
        grid = a.shape[1:]
        return a.reshape(a.shape[0], -1), grid
    return a.reshape(-1), a.shape


class InitData(dict):
    """Mapping with the ``.files`` attribute of an ``.npz`` so that ``driver.columns_from_init`` treats both alike."""
    grid: tuple = ()

    @property
    def files(self):
        return list(self.keys())


STATIC_KEYS = ("lat", "lon", "lutyp", "sltyp", "slptyp", "isc", "shdmax", "tbot")   # what a geo_em-style file may supply


def _read_into(out: "InitData", path: str, keys) -> None:
    with _open(path) as nc:
        for key in keys:
            if key in out:
                continue
            name = next((n for n in INIT_ALIASES[key] if n in nc.variables), None)
            if name is None:
                continue
            a, dims = _squeeze_time(nc.variables[name])
            a, grid = _flatten(a, dims, key in _SOIL_KEYS)
            if key in _PERCENT_KEYS and name in ("VEGFRA", "GVFMAX"):
                a = a.astype(f32) / f32(100.0)
            out[key] = a
            if key == "lat":
                out.grid = tuple(grid)


def read_init(path: str, static_path: str | None = None) -> InitData:
    """The initialisation file; fields it lacks among ``STATIC_KEYS`` are taken from the namelist's
    ``static_parameter_file`` (``geo_em.d01.nc``-style: XLAT XLONG LU_INDEX ...) when that is a NetCDF-3 file on the same grid."""
    out = InitData()
    _read_into(out, path, INIT_ALIASES)
    if static_path and os.path.isfile(static_path) and is_netcdf(static_path):
        _read_into(out, static_path, STATIC_KEYS)
        n = out["lat"].size if "lat" in out else None
        for key in STATIC_KEYS:
            if key in out and n is not None and out[key].size != n:
                raise ValueError(f"ERR: {static_path}: {key} has {out[key].size} cells, {path} has {n}")
    for need in ("lat", "lon", "lutyp", "sltyp", "tsk", "stc", "smc"):
        if need not in out:
            raise ValueError(f"ERR: {path} has none of {INIT_ALIASES[need]} (needed for {need})")
    return out


_RECORD_LAYOUTS: dict = {}     # (header bytes, ncol) -> [(offset of the variable's data, dtype string)] in forcing.RECORD_VARS order


def _record_layout(path: str, ncol: int):
    """Parse one record file with scipy and remember where its eight variables start.  Files written by the same program
    for the same grid have byte-identical headers, so the layout of the first record serves every later one whose header
    compares equal (checked per file); anything else is parsed on its own."""
    from scipy.io import netcdf_file
    nc = netcdf_file(path, "r", mmap=True, version=2)
    try:
        base = nc._mm_buf.__array_interface__["data"][0]
        lay = []
        for names in FORCING_ALIASES:
            name = next((n for n in names if n in nc.variables), None)
            if name is None:
                raise ValueError(f"ERR: {path} has none of {names}")
            v = nc.variables[name]
            if int(np.prod(v.shape)) != ncol:
                raise ValueError(f"{path}: {name} has {int(np.prod(v.shape))} cells, the run has {ncol} columns")
            if v.isrec:
                return None            # record variables are interleaved: leave those files to scipy
            d = v.data
            lay.append((d.__array_interface__["data"][0] - base, d.dtype.str))
            del d, v
    except AttributeError:             # scipy internals moved: no fast path
        return None
    finally:
        import warnings
        with warnings.catch_warnings():      # after an error a view of the map may still be alive: scipy warns and leaves it
            warnings.simplefilter("ignore")  # to the garbage collector, which is fine for a read-only map
            try:
                nc.close()
            except Exception:
                pass
    return lay


def read_record(path: str, ncol: int, columns: slice | None = None) -> np.ndarray:
    """One forcing record as float32 [8][ncol] in the order of ``forcing.RECORD_VARS``; with ``columns`` only that block of
    the flattened grid, [8][columns.stop - columns.start] (a rank of a multi-GPU run reads its own columns: the file is
    memory-mapped and only the pages of the block are touched).  The header is parsed once per file layout, not per file."""
    c0, c1 = (0, ncol) if columns is None else (columns.start, columns.stop)
    rec = np.empty((len(FORCING_ALIASES), c1 - c0), f32)
    with open(path, "rb") as fh:
        head = fh.read(4096)
    lay = None
    for (hdr, n), cand in _RECORD_LAYOUTS.items():
        if n == ncol and head.startswith(hdr):
            lay = cand
            break
    if lay is None:
        lay = _record_layout(path, ncol)
        if lay is not None:
            hlen = min(off for off, _ in lay)
            if hlen <= len(head) and len(_RECORD_LAYOUTS) < 16:
                _RECORD_LAYOUTS[(head[:hlen], ncol)] = lay     # the cell count was checked against this ncol
    if lay is not None:
        fd = os.open(path, os.O_RDONLY)      # positioned reads of the block: faster than faulting a map in page by page
        try:
            size = os.fstat(fd).st_size
            for i, (off, dts) in enumerate(lay):
                item = np.dtype(dts).itemsize
                if off + ncol * item > size:
                    raise ValueError(f"{path}: truncated (variable {FORCING_ALIASES[i][0]} ends beyond the file)")
                if item == 4 and np.dtype(dts).kind == "f":      # straight into the row, byte order fixed in place
                    got = os.preadv(fd, [memoryview(rec[i]).cast("B")], off + c0 * item)
                    if got != (c1 - c0) * item:
                        raise ValueError(f"{path}: short read of {FORCING_ALIASES[i][0]}")
                    if np.dtype(dts).byteorder == ">" or (np.dtype(dts).byteorder == "=" and sys.byteorder == "big"):
                        rec[i].view(np.uint32).byteswap(inplace=True)
                else:
                    raw = os.pread(fd, (c1 - c0) * item, off + c0 * item)
                    if len(raw) != (c1 - c0) * item:
                        raise ValueError(f"{path}: short read of {FORCING_ALIASES[i][0]}")
                    rec[i] = np.frombuffer(raw, dtype=dts)
        finally:
            os.close(fd)
        return rec
    from scipy.io import netcdf_file                      # general path: let scipy interpret the file
    nc = netcdf_file(path, "r", mmap=False, version=2)
    try:
        for i, names in enumerate(FORCING_ALIASES):
            name = next((n for n in names if n in nc.variables), None)
            if name is None:
                raise ValueError(f"ERR: {path} has none of {names}")
            a, _ = _squeeze_time(nc.variables[name])
            if a.size != ncol:
                raise ValueError(f"{path}: {name} has {a.size} cells, the run has {ncol} columns")
            rec[i] = a.reshape(-1)[c0:c1]
    finally:
        nc.close()
    return rec


def _grid_dims(nc, grid, ncol):
    if len(grid) == 2 and grid[0] * grid[1] == ncol:
        nc.createDimension("south_north", grid[0])
        nc.createDimension("west_east", grid[1])
        return ("Time", "south_north", "west_east"), (1, grid[0], grid[1])
    nc.createDimension("ncol", ncol)
    return ("Time", "ncol"), (1, ncol)


def write_planes(path: str, planes: np.ndarray, names, when: dt.datetime, grid=(), int_planes=(), attrs=None) -> None:
    """One float32 variable per plane (int32 for ``int_planes``, stored from the bit pattern) on the run's grid."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    ncol = planes.shape[1]
    with _open(path, "w") as nc:
        nc.createDimension("Time", 1)
        nc.createDimension("DateStrLen", 19)
        dims, shape = _grid_dims(nc, tuple(grid), ncol)
        t = nc.createVariable("Times", "S1", ("Time", "DateStrLen"))
        t[0, :] = np.frombuffer(when.strftime("%Y-%m-%d_%H:%M:%S").encode(), dtype="S1")
        for i, name in enumerate(names):
            safe = name.replace("+", "p").replace("-", "m")      # STC+1 -> STCp1, STC-2 -> STCm2
            if name in int_planes:
                v = nc.createVariable(safe, "i4", dims)
                v[:] = planes[i].view(np.int32).reshape(shape)
            else:
                v = nc.createVariable(safe, "f4", dims)
                v[:] = planes[i].reshape(shape)
            v.plane = name.encode()
        nc.valid_time = when.strftime("%Y%m%d%H%M").encode()
        for k, val in (attrs or {}).items():
            setattr(nc, k, val)


def read_planes(path: str, names, int_planes=()):
    """Inverse of :func:`write_planes`: ([nplane][ncol] float32 with int planes as bit patterns, time, attributes)."""
    with _open(path) as nc:
        first = names[0].replace("+", "p").replace("-", "m")
        ncol = int(np.prod(nc.variables[first].shape))
        out = np.empty((len(names), ncol), f32)
        for i, name in enumerate(names):
            safe = name.replace("+", "p").replace("-", "m")
            if safe not in nc.variables:
                raise ValueError(f"{path}: plane {name} is missing - written by a different build?")
            a = np.array(nc.variables[safe].data).reshape(-1)
            out[i] = a.astype(np.int32).view(f32) if name in int_planes else a.astype(f32)
        when = dt.datetime.strptime(bytes(nc.valid_time).decode(), "%Y%m%d%H%M")
        attrs = {k: getattr(nc, k) for k in ("model_step",) if hasattr(nc, k)}
    return out, when, attrs


def write_restart(path: str, state: np.ndarray, when: dt.datetime, step: int, grid=()) -> None:
    write_planes(path, state, L.STATE, when, grid, int_planes=("ISNOW",), attrs={"model_step": np.int32(step)})


def read_restart(path: str):
    st, when, attrs = read_planes(path, L.STATE, int_planes=("ISNOW",))
    return st, when, int(np.asarray(attrs.get("model_step", 0)).reshape(-1)[0])


def write_output(path: str, diag: np.ndarray, when: dt.datetime, grid=()) -> None:
    write_planes(path, diag, L.DIAG, when, grid, int_planes=("ITER", "FLAGS"))
