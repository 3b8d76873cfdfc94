"""Forcing ingest for the namelist-driven driver: record files at ``input_frequency``, time
interpolation to ``interval_seconds`` and the cosine of the solar zenith angle -- none of which the reference
implements (it only carries the config fields, offline/noahmp_config.py:81-86,98-102).

Host side here = file I/O and the handful of per-step scalars; the per-column arithmetic runs on the GPU
(``noahmp_prepare_forcing``).  Record files: ``<input_directory>/<YYYYMMDDHHMM>.npy``, float32 ``[8][ncol]`` in
the order SFCTMP SFCPRS UU VV Q2 SOLDN LWDN PRCP, or NetCDF-3 records ``<YYYYMMDDHH>.LDASIN_DOMAIN1`` /
``<YYYYMMDDHHMM>.nc`` with the HRLDAS variable names (``ncio.py``)."""
from __future__ import annotations

import calendar
import datetime as dt
import math
import os
import re

import numpy as np

RECORD_VARS = ["SFCTMP", "SFCPRS", "UU", "VV", "Q2", "SOLDN", "LWDN", "PRCP"]
_UNITS = {"second": 1, "minute": 60, "hour": 3600, "day": 86400}


def parse_frequency(text) -> tuple[int, str]:
    """'1 hour' -> (3600, 'seconds'); '1 month' -> (1, 'months').  Accepts plural units and bare integers (seconds)."""
    if isinstance(text, (int, float)):
        return int(text), "seconds"
    m = re.fullmatch(r"\s*(\d+)\s*([A-Za-z]+)\s*", str(text))
    if not m:
        raise ValueError(f"cannot parse frequency {text!r}")
    n, unit = int(m.group(1)), m.group(2).lower().rstrip("s")
    if unit == "month":
        return n, "months"
    if unit == "year":
        return 12 * n, "months"
    if unit not in _UNITS:
        raise ValueError(f"unknown time unit in {text!r}")
    return n * _UNITS[unit], "seconds"


def advance(t: dt.datetime, freq: tuple[int, str]) -> dt.datetime:
    n, kind = freq
    if kind == "seconds":
        return t + dt.timedelta(seconds=n)
    month = t.month - 1 + n
    year, month = t.year + month // 12, month % 12 + 1
    return t.replace(year=year, month=month, day=min(t.day, calendar.monthrange(year, month)[1]))


def is_boundary(t: dt.datetime, origin: dt.datetime, freq: tuple[int, str]) -> bool:
    """True when ``t`` is ``origin`` plus a whole number of ``freq`` periods."""
    n, kind = freq
    if kind == "seconds":
        return (t - origin).total_seconds() % n == 0
    months = (t.year - origin.year) * 12 + (t.month - origin.month)
    return months % n == 0 and (t.day, t.hour, t.minute, t.second) == (origin.day, origin.hour, origin.minute, origin.second)


def stamp(t: dt.datetime) -> str:
    return t.strftime("%Y%m%d%H%M")


def julian_of(t: dt.datetime) -> float:
    """JULIAN of noahmp_sflx (func:127): day of year as a float, 0.0 at 00:00 on 1 January."""
    return (t.timetuple().tm_yday - 1) + (t.hour * 3600 + t.minute * 60 + t.second) / 86400.0


def yearlen_of(t: dt.datetime) -> int:
    return 366 if calendar.isleap(t.year) else 365


def solar_scalars(t: dt.datetime) -> np.ndarray:
    """sin(decl), cos(decl), cos(H0), sin(H0) for UTC time ``t``; H0 = hour angle at longitude 0.  Declination
    by the cosine formula (23.44 deg amplitude, solstice on day 172); no equation of time."""
    doy = julian_of(t) + 1.0
    decl = math.radians(23.44) * math.cos(2.0 * math.pi * (doy - 172.0) / yearlen_of(t))
    h0 = 2.0 * math.pi * ((t.hour + t.minute / 60.0 + t.second / 3600.0) - 12.0) / 24.0
    return np.asarray([math.sin(decl), math.cos(decl), math.cos(h0), math.sin(h0)], np.float32)


def geo_planes(lat_rad, lon_rad) -> np.ndarray:
    """[4][ncol] fp32: sin(lat) cos(lat) sin(lon) cos(lon), evaluated in fp64 and rounded once."""
    lat, lon = np.asarray(lat_rad, np.float64), np.asarray(lon_rad, np.float64)
    return np.stack([np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)]).astype(np.float32)


class RecordReader:
    """Forcing records on disk, cached two at a time on the device."""

    def __init__(self, directory: str, freq, origin: dt.datetime, ncol: int, device):
        self.dir, self.freq, self.origin, self.ncol, self.device = directory, parse_frequency(freq), origin, ncol, device
        if self.freq[1] != "seconds":
            raise ValueError("input_frequency must be a fixed number of seconds")
        self._cache: dict = {}

    def path(self, t: dt.datetime) -> str:
        """The record of time ``t``: ``.npy`` first, then the netCDF names; the ``.npy`` name when none exists."""
        names = [stamp(t) + ".npy", stamp(t) + ".nc", t.strftime("%Y%m%d%H") + ".LDASIN_DOMAIN1", stamp(t) + ".LDASIN_DOMAIN1"]
        for n in names:
            if os.path.isfile(os.path.join(self.dir, n)):
                return os.path.join(self.dir, n)
        return os.path.join(self.dir, names[0])

    def bracket(self, t: dt.datetime):
        """(t0, t1, weight) with t0 <= t < t1 on the record grid."""
        n = self.freq[0]
        k = int((t - self.origin).total_seconds() // n)
        t0 = self.origin + dt.timedelta(seconds=k * n)
        w = (t - t0).total_seconds() / n
        # on a record time the weight is 0 and the later record is not needed: return t0 twice, so that the last step of
        # a run does not ask for a record dated after the end of the forcing period
        t1 = t0 if w == 0.0 else t0 + dt.timedelta(seconds=n)
        return t0, t1, w

    def load(self, t: dt.datetime, columns: slice | None = None) -> np.ndarray:
        """The record of time ``t`` as a host array float32 [8][ncol] (no caching); with ``columns`` only that block of
        columns, read through a memory map so that a rank of a multi-GPU run touches its own part of the file only."""
        p = self.path(t)
        if not os.path.isfile(p):
            raise FileNotFoundError(f"ERR: Unable to find forcing record {p}")
        if p.endswith(".npy"):
            a = np.load(p, mmap_mode="r" if columns is not None else None)
            if a.shape != (len(RECORD_VARS), self.ncol) or a.dtype != np.float32:
                raise ValueError(f"{p}: expected float32 [{len(RECORD_VARS)}][{self.ncol}], got {a.dtype} {a.shape}")
            return np.array(a[:, columns]) if columns is not None else a      # a copy: the map is read-only
        from . import ncio
        return ncio.read_record(p, self.ncol, columns)

    def record(self, t: dt.datetime):
        import torch
        key = stamp(t)
        if key not in self._cache:
            a = self.load(t)
            if len(self._cache) >= 2:
                self._cache.pop(sorted(self._cache)[0])
            self._cache[key] = torch.from_numpy(a).to(self.device)
        return self._cache[key]


class RecordPrefetcher:
    """Reads the records of a run ahead of the time loop, in time order, on a small pool of threads, into page-locked host
    buffers (so that the copy to the device is asynchronous too).  The time loop takes them in the same order; anything
    unexpected (a record asked for out of order, a missing file) makes :meth:`take` return None and the caller reads the
    record itself, with the reader's own error messages."""

    def __init__(self, reader: RecordReader, columns: slice, first: dt.datetime, last: dt.datetime, depth: int = 16,
                 pinned: bool = True, threads: int = 4):
        import collections
        from concurrent.futures import ThreadPoolExecutor
        import torch
        self.reader, self.columns, self.depth = reader, columns, depth
        n = columns.stop - columns.start
        self.free = [torch.empty((len(RECORD_VARS), n), dtype=torch.float32, pin_memory=pinned) for _ in range(depth)]
        self.pending: "collections.deque" = collections.deque()     # (stamp, future, buffer) in time order
        self.pool = ThreadPoolExecutor(max_workers=threads)
        self.busy_s = 0.0
        self._dead = False
        step = dt.timedelta(seconds=reader.freq[0])
        self._times = iter([first + k * step for k in range(int((last - first) / step) + 1)])
        self._fill()

    def _load(self, t, buf):
        import time
        tic = time.perf_counter()
        buf.numpy()[...] = self.reader.load(t, self.columns)
        self.busy_s += time.perf_counter() - tic      # summed over the pool's threads
        return buf

    def _fill(self) -> None:
        while self.free and not self._dead:
            t = next(self._times, None)
            if t is None:
                return
            buf = self.free.pop()
            self.pending.append((stamp(t), self.pool.submit(self._load, t, buf), buf))

    def take(self, key: str):
        """The buffer holding record ``key`` (to be handed back with :meth:`give_back`), or None."""
        while not self._dead and self.pending:
            k, fut, buf = self.pending.popleft()
            if k > key:                # the loop asked for an earlier record than the readers have reached: out of order
                self.pending.appendleft((k, fut, buf))
                self._dead = True
                return None
            try:
                fut.result()
            except Exception:          # missing / malformed file: the time loop reads it itself and reports
                self._dead = True
                return None
            if k == key:
                self._fill()
                return buf
            self.free.append(buf)      # a record the loop skipped
            self._fill()
        return None

    def give_back(self, buf) -> None:
        self.free.append(buf)
        self._fill()

    def close(self) -> None:
        self._dead = True
        self.pool.shutdown(wait=True, cancel_futures=True)


class RecordRing:
    """Device ring of forcing records for records-driven launches (``noahmp_run_records``): a record is read from disk and
    copied to the device once, however many model steps and launches it serves.  With a :class:`RecordPrefetcher` the file
    reads run ahead on a background thread and the copies are asynchronous (page-locked buffers)."""

    def __init__(self, reader: RecordReader, slots: int, columns: slice, device, prefetch: "RecordPrefetcher | None" = None):
        import torch
        self.reader, self.columns = reader, columns
        n = columns.stop - columns.start
        self.ring = torch.empty((slots, len(RECORD_VARS), n), dtype=torch.float32, device=device)
        self.held: dict = {}      # stamp -> slot
        self.loads = 0
        self.prefetch = prefetch
        self.prefetched = 0
        self._inflight: list = []  # (event, pinned buffer) of copies not yet known to be complete

    def slots_for(self, times) -> dict:
        """Make every record time in ``times`` resident; returns {time: slot}.  Slots of records this launch does not need
        are recycled (stream order keeps earlier launches safe: copies and launches share the stream)."""
        import torch
        need = {stamp(t): t for t in times}
        if len(need) > self.ring.shape[0]:
            raise ValueError("a launch needs more records than the ring holds")
        free = [s for s in range(self.ring.shape[0]) if s not in self.held.values()]
        for key in [k for k in self.held if k not in need]:
            free.append(self.held.pop(key))
        for ev, buf in [x for x in self._inflight if x[0].query()]:
            self.prefetch.give_back(buf)
        self._inflight = [x for x in self._inflight if not x[0].query()]
        for key in sorted(need):          # time order: the order the prefetcher reads in
            if key not in self.held:
                slot = free.pop()
                buf = self.prefetch.take(key) if self.prefetch is not None else None
                if buf is not None:
                    self.ring[slot].copy_(buf, non_blocking=True)
                    if self.ring.is_cuda:
                        ev = torch.cuda.Event()
                        ev.record()
                        self._inflight.append((ev, buf))
                        while len(self._inflight) > max(1, self.prefetch.depth - 3):   # never starve the reader of buffers
                            ev0, buf0 = self._inflight.pop(0)
                            ev0.synchronize()
                            self.prefetch.give_back(buf0)
                    else:
                        self.prefetch.give_back(buf)
                    self.prefetched += 1
                else:
                    self.ring[slot].copy_(torch.from_numpy(self.reader.load(need[key], self.columns)), non_blocking=False)
                self.held[key] = slot
                self.loads += 1
        return {t: self.held[stamp(t)] for t in times}
