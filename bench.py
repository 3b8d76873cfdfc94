#!/usr/bin/env python3
"""Benchmark of the Noah-MP column time step on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle restatement)

metric   cell-timesteps/s  = columns x model-steps advanced per second, whole job (all ranks)
step     one launch of the step kernel advancing EVERY column of the workload by the workload's
         `model_steps` time steps (time-fused in one kernel; forcing of every model step is read from HBM)
workload BASELINE.json configs[3]: CONUS 1 km synthetic grid, 17e6 columns x 24 hourly steps, mixed
         soil/vegetation types, block-partitioned over the ranks (strong scaling, no data-path collective;
         one NCCL all-reduce of the 16 budget sums after the timed region)
value    inputs resident in HBM; CUDA events on the launching stream; max over ranks
e2e      the same work through the C-ABI host entry point `noahmp_advance_host`: state resident on the device,
         every model step's forcing copied from pinned HOST memory (overlapped with compute) and the budget sums
         read back, all inside the timed region; `e2e_state_roundtrip` = `noahmp_run_host`, which also moves the
         whole state H2D and D2H
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from noahmp_b200 import layout as L, synth  # noqa: E402
from noahmp_b200.parallel import allreduce_budget, budget_means, device_for_rank, divergence_order, shard  # noqa: E402
from noahmp_b200.tables import ParamTables  # noqa: E402

B_ALG = 404  # algorithmic bytes per cell-step (SURVEY 8d / DESIGN.md section 4)
METRIC = "cell_timesteps_per_sec"
UNIT = "cell-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="conus1km")
    ap.add_argument("--ncol", type=int, default=None, help="override the workload's TOTAL column count")
    ap.add_argument("--model-steps", type=int, default=None, help="override model time steps per launch")
    ap.add_argument("--variant", default="fast", choices=["parity", "fast", "strict"])
    ap.add_argument("--no-sort-columns", action="store_true", help="keep the generator's random column order")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-variants", action="store_true", help="skip timing the parity/strict builds")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the other BASELINE configs (N = 1 default line)")
    ap.add_argument("--e2e-steps-per-launch", type=int, default=1)
    ap.add_argument("--pack-threads", type=int, default=0, help="host threads per rank that pack record planes (0: this rank's share of the cores, at most 16)")
    ap.add_argument("--pack-sweep", default="", help="comma-separated thread counts: time the packed e2e leg with each (diagnostic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-cols", type=int, default=1_000_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _capture():
    """The committed ncu capture of the step kernel on the benchmark workload: this round's if there is one."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                t = json.load(fh)
            t["file"] = "profiles/" + name
            return t
        except Exception:
            continue
    return None


def ncu_traffic(nloc, nsteps, variant):
    """(DRAM bytes per launch, where the figure comes from): read from the committed `ncu --set full` capture when it is of
    this exact launch (same columns, fused steps and build) - it is a citation of that capture, not a live measurement."""
    t = _capture()
    if t and t["columns"] == nloc and t["model_steps"] == nsteps and t["variant"] == variant:
        return t["traffic_bytes_per_launch"], f"{t['file']} ({t.get('capture', 'ncu --set full')}); not measured in this run"
    return None, "no committed capture of this exact launch"


def issue_roofline(cell_steps_per_s, variant, sm_count, sm_mhz):
    """The bound the kernel actually runs against: one warp-instruction per scheduler per clock (4 schedulers per SM).
    Warp-instructions per cell-step come from the committed ncu capture of this workload (profiles/r01_traffic.json)."""
    try:
        t = _capture()
        if t is None or t["variant"] != variant or not sm_mhz:
            return None
        per = t["warp_inst"] / (t["columns"] * t["model_steps"])
        peak = sm_count * 4 * sm_mhz * 1e6
        return {"bound": "issue", "warp_inst_per_cell_step": per, "achieved": cell_steps_per_s * per, "peak": peak,
                "unit": "warp-inst/s", "frac": cell_steps_per_s * per / peak,
                "source": f"instruction count of {t['file']} (ncu), SM clock sampled during the timed region"}
    except Exception:
        return None


def build_workload(args):
    """The workload of the generators, re-ordered by divergence class.  `w.order[j]` = generator id of the column at position
    j: the forcing generator hashes that id, so every column keeps its own forcing whatever the order and the rank count."""
    w = synth.make_workload(args.workload, ncol=args.ncol, nsteps=args.model_steps)
    w.order = np.arange(w.ncol, dtype=np.int64)
    if not args.no_sort_columns:
        # divergence-class ordering, balanced over the ranks; done once at set-up
        world = int(os.environ.get("WORLD_SIZE", "1"))
        order = divergence_order(w.static_i, w.static_f[L.C["LAT"]], world)
        w.state = np.ascontiguousarray(w.state[:, order])
        w.static_f = np.ascontiguousarray(w.static_f[:, order])
        w.static_i = np.ascontiguousarray(w.static_i[order])
        w.sincos = np.ascontiguousarray(w.sincos[:, order])
        w.order = np.ascontiguousarray(order.astype(np.int64))
    return w


def slice_workload(w, c0, n):
    import copy
    s = copy.copy(w)
    s.ncol = n
    s.order = np.ascontiguousarray(w.order[c0:c0 + n])
    s.state = np.ascontiguousarray(w.state[:, c0:c0 + n])
    s.static_f = np.ascontiguousarray(w.static_f[:, c0:c0 + n])
    s.static_i = np.ascontiguousarray(w.static_i[c0:c0 + n])
    s.sincos = np.ascontiguousarray(w.sincos[:, c0:c0 + n])
    return s


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which the CPU legs must not obey)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def sample_workload(w, n):
    """Every k-th column of the (class-sorted) workload: a representative sample with the same class mix."""
    import copy
    n = min(n, w.ncol)
    idx = (np.arange(n, dtype=np.int64) * w.ncol) // n
    s = copy.copy(w)
    s.ncol = n
    s.state = np.ascontiguousarray(w.state[:, idx])
    s.static_f = np.ascontiguousarray(w.static_f[:, idx])
    s.static_i = np.ascontiguousarray(w.static_i[idx])
    s.sincos = np.ascontiguousarray(w.sincos[:, idx])
    s.order = np.ascontiguousarray(w.order[idx])
    return s, idx


def config_of(w, args, extra=None):
    cfg = {"workload": w.name, "columns": w.ncol, "model_steps_per_step": w.nsteps, "dt_s": w.dt, "options": w.opts,
           "column_order": "random (generator)" if args.no_sort_columns else "land-use type, latitude band; dealt round-robin to ranks"}
    cfg.update(extra or {})
    return cfg


# ------------------------------------------------------------------------------------------------
def cpu_baseline(w, sample_cols: int, fast: bool = True, repeats: int = 1) -> dict:
    """The reference's CPU path (C++ restatement of core/module_noahmp_func.f90; the Fortran cannot be
    compiled here) on all host cores, on the first `sample_cols` columns x all model steps."""
    from oracle.oracle import Oracle
    ws, idx = sample_workload(w, sample_cols)
    n = ws.ncol
    forcing = synth.forcing_all(ws, cols=ws.order)
    o = Oracle(ParamTables.default(), w.opts, fast="fast" if fast else "parity")
    best = None
    for _ in range(repeats):
        st = ws.state.copy()
        t0 = time.perf_counter()
        o.run_workload(ws, forcing, state=st, nthreads=host_threads())
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": n * w.nsteps / best, "unit": UNIT, "cores": host_threads(), "kind": "port",
            "sample": f"every {w.ncol // n}-th column ({n} columns) x {w.nsteps} model steps of {w.name} "
                      f"({'-O3 AVX2+FMA' if fast else '-O2 no-FMA'} build of the C++ restatement, OpenMP static, {best:.2f} s)"}


def cpu_leg_of(name: str, variant: str, local: int, sample_cols: int = 20_000, max_steps: int = 48) -> dict:
    """CPU figure and iteration-count check of one BASELINE config on a bounded sample (part of the cpu_baseline leg: the
    oracle is the thing timed, and the checker of the counters).  (1) the -O3 port on all host cores on `sample_cols` columns
    x min(steps, max_steps) model steps; (2) north star "identical iteration counts in the energy-balance solver": from the
    SAME state (the plain oracle's after all but the last of those steps) one more step by the plain oracle and by the GPU
    build, packed counters compared per column - canopy Newton loop (func:2870-2875), stomata bisections (func:3830-3842),
    soil sub-steps (func:5971-5978).  The exit tests sit on fp32 values that differ in the last ulps between any two libm
    implementations, so the yardstick is printed beside it: the same comparison between two CPU builds of the oracle."""
    import torch
    from oracle.oracle import Oracle
    from noahmp_b200.engine import NoahMP
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import accuracy as A
    ncol_full, nsteps_full = {"single": (1, 96), "uniform100k": (100_000, 48), "global025": (280_000, 8760),
                              "conus1km": (17_000_000, 24), "snow1m": (1_000_000, 720)}[name]
    n, nsteps = min(sample_cols, ncol_full), min(max_steps, nsteps_full)
    w = synth.make_workload(name, ncol=n, nsteps=nsteps)
    forcing = synth.forcing_all(w)
    tables = ParamTables.default()
    st = w.state.copy()
    t0 = time.perf_counter()
    Oracle(tables, w.opts, fast="fast").run_workload(w, forcing, state=st, nthreads=host_threads())
    dt = time.perf_counter() - t0
    out = {"value": n * nsteps / dt, "unit": UNIT, "cores": host_threads(), "kind": "port",
           "sample": f"{n} columns x {nsteps} model steps of {name} (-O3 AVX2+FMA build, OpenMP static, {dt:.2f} s)"}
    if n < 64 or nsteps < 2:
        return out
    kw = dict(dt=w.dt, zsoil=synth.ZSOIL, yearlen=w.yearlen, zref=w.zref, co2_frac=w.co2_frac, o2_frac=w.o2_frac, foln=w.foln)
    s0 = w.state.copy()
    # the compared step is the sunniest one of the sample (the stomatal bisections only run by day)
    m = int(np.argmax(forcing[1:, L.FORCING.index("COSZ")].mean(axis=1))) + 1
    Oracle(tables, w.opts, fast=False).run(s0, np.ascontiguousarray(forcing[:m]), w.static_f, w.static_i, julian0=w.julian0,
                                           nthreads=host_threads(), **kw)
    jlast = w.julian0 + m * w.dt / 86400.0
    last = np.ascontiguousarray(forcing[m:m + 1])

    def cpu_step(kind):
        s, d = s0.copy(), np.zeros((L.NDIAG, n), np.float32)
        Oracle(tables, w.opts, fast=kind).run(s, last, w.static_f, w.static_i, julian0=jlast, diag=d, nthreads=host_threads(), **kw)
        return d[L.D["ITER"]]

    it_ref = cpu_step(False)
    dev = torch.device("cuda", local)
    eng = NoahMP(tables, w.opts, device=local, fast=variant)
    dg = torch.zeros((L.NDIAG, n), dtype=torch.float32, device=dev)
    eng.run(eng.run_params(w.dt, synth.ZSOIL, jlast, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln), torch.from_numpy(s0.copy()).to(dev),
            torch.from_numpy(last).to(dev), torch.from_numpy(w.static_f).to(dev), torch.from_numpy(w.static_i).to(dev), dg, None)
    torch.cuda.synchronize()
    it_gpu = dg[L.D["ITER"]].cpu().numpy()
    eng.close()
    out["iteration_count_mismatch"] = {"gpu_vs_oracle": A.iteration_mismatch(it_gpu, it_ref),
                                       "two_cpu_builds_of_the_oracle": A.iteration_mismatch(cpu_step("fast"), it_ref),
                                       "columns": n, "after_steps": m,
                                       "what": "fraction of columns whose counter differs on one step taken from identical states"}
    return out


def translated_reference_leg(w, sample_cols: int = 200_000):
    """The reference's own statements (mechanical Fortran -> C++ translation, oracle/_ref, -O2, OpenMP over columns) timed on a
    smaller sample.  It is about half as fast as the hand-written port, so the port stays the baseline the GPU is set against
    (the conservative choice); this figure is reported beside it.  None when oracle/_ref was not built (no /root/reference)."""
    try:
        from oracle import refrun
        if not (os.path.exists(refrun.LIB) and os.path.exists(refrun.SIG)):
            return None
        ws, idx = sample_workload(w, sample_cols)
        forcing = synth.forcing_all(ws, cols=ws.order)
        ref = refrun.RefModel(ParamTables.default(), w.opts)
        st = ws.state.copy()
        t0 = time.perf_counter()
        fatal = ref.run_native(st, forcing, ws.static_f, ws.static_i, w.dt, synth.ZSOIL, w.julian0, w.yearlen, w.zref, w.co2_frac,
                               w.o2_frac, w.foln, nthreads=host_threads())
        dt = time.perf_counter() - t0
        return {"value": ws.ncol * w.nsteps / dt, "unit": UNIT, "cores": host_threads(), "kind": "reference (translated)",
                "sample": f"{ws.ncol} columns x {w.nsteps} model steps, {dt:.2f} s, {int(fatal.sum())} columns aborted"}
    except Exception as exc:  # the baseline legs must never take the bench line down
        return {"error": f"{type(exc).__name__}: {exc}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = build_workload(args)
    from oracle.oracle import Oracle
    ws, idx = sample_workload(w, args.cpu_sample_cols)
    n = ws.ncol
    forcing = synth.forcing_all(ws, cols=ws.order)
    o = Oracle(ParamTables.default(), w.opts, fast="fast")
    times = []
    for it in range(args.warmup + args.steps):
        st = ws.state.copy()
        t0 = time.perf_counter()
        o.run_workload(ws, forcing, state=st, nthreads=host_threads())
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = n * w.nsteps * len(times) / total
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(w, args, {"sampled_columns": n}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads(), "kind": "port",
                             "sample": f"every {w.ncol // n}-th column ({n} columns) x {w.nsteps} model steps per step (-O3 AVX2+FMA "
                                       "build of the C++ restatement; the reference Fortran cannot be compiled in this image)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    tr = translated_reference_leg(w)
    if tr:
        line["cpu_baseline"]["translated_reference"] = tr
    emit(line)


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local: int) -> dict:
    """Pin this rank to the CPUs next to its GPU (sysfs: the PCI device's local_cpulist / numa_node) BEFORE any pinned
    buffer is allocated, so that the page-locked forcing records are first touched on the GPU's own NUMA node and the H2D
    copies do not cross the socket interconnect.  No-op (and says so) where the topology is not exposed."""
    import torch
    info = {"applied": False}
    try:
        props = torch.cuda.get_device_properties(local)
        bus = f"{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        info.update({"pci": bus, "numa_node": node, "local_cpulist": cpulist})
        cpus = set()
        for part in cpulist.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if node >= 0 and allowed and len(allowed) < len(os.sched_getaffinity(0)) and world > 1:
            os.sched_setaffinity(0, allowed)
            info["applied"] = True
            info["cpus_bound"] = len(allowed)
    except Exception as exc:  # topology not exposed (container, single node): leave the scheduler alone
        info["note"] = f"{type(exc).__name__}: {exc}"
    return info


def h2d_probe(dev, world, nbytes=1 << 30, reps=6) -> dict:
    """Host->device copy bandwidth of pinned memory with ALL ranks copying at once and no compute: the ceiling of any
    end-to-end figure that streams forcing (a model step needs 32 B per column over this link)."""
    import torch
    import torch.distributed as dist
    h = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    h.fill_(1.0)
    d = torch.empty_like(h, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    gbs = nbytes * reps / float(el.item()) / 1e9
    # each rank's own rate while ALL ranks keep copying (fixed duration, not fixed size): the links of a node need not be equal
    chunk = h[: (1 << 29) // 4]
    dchunk = d[: (1 << 29) // 4]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, moved = time.perf_counter(), 0
    while time.perf_counter() - t0 < 0.8:
        dchunk.copy_(chunk, non_blocking=True)
        torch.cuda.synchronize()
        moved += chunk.numel() * 4
    mine = torch.tensor([moved / (time.perf_counter() - t0) / 1e9], dtype=torch.float64, device=dev)
    each = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(each, mine)
    else:
        each = [mine]
    rates = [float(x.item()) for x in each]
    del h, d
    return {"h2d_gbs_per_gpu": gbs, "h2d_gbs_all_gpus": gbs * world, "ranks_copying": world, "bytes_per_copy": nbytes, "copies": reps,
            "e2e_ceiling_cell_steps_per_s": gbs * world * 1e9 / 32.0,
            "h2d_gbs_each_rank_while_all_copy": [round(r, 1) for r in rates], "h2d_gbs_sum_while_all_copy": sum(rates),
            "e2e_ceiling_link_weighted_cell_steps_per_s": sum(rates) * 1e9 / 32.0,
            "note": "pinned host memory, all ranks at once, slowest rank's time; a model step streams 8 fp32 planes = 32 B per column"}


def twin_sample(name, order, c0, nloc):
    """Columns of the committed fp64-twin fixture (tests/golden/twin_<name>.npz) that this rank owns: (fixture, selector into the
    fixture, local column positions).  `order` maps sorted position -> generator column id."""
    path = os.path.join(ROOT, "tests", "golden", f"twin_{name}.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    ids = g["columns"]
    if order is None:
        pos = ids
    else:
        inv = np.empty(order.size, np.int64)
        inv[order] = np.arange(order.size)
        pos = inv[ids]
    mine = (pos >= c0) & (pos < c0 + nloc)
    return g, np.where(mine)[0], pos[mine] - c0


def parity_of(state_planes, fixture, sel):
    """Distance of the sampled columns from the fp64 twin of the reference, beside the fp32 reference's own (tools/accuracy.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import accuracy as A
    cmp = A.compare(state_planes[:L.NSTATE_CORE], fixture["ref32"][:, sel], fixture["ref64"][:, sel])
    out = A.summary(cmp)
    out["source"] = "sampled columns of this run against tests/golden/twin_*.npz (translated reference, fp32 and fp64, full step count)"
    return out


def run_other_workload(name, nsteps, twin_name, variant, dev, stream, peak, sm_count, clocks_mhz):
    """One more BASELINE config on this GPU, full size and full length: forcing generated on the device 24 steps at a time, the
    step launches timed with CUDA events (forcing generation outside the events), sampled columns set against the fp64 twin."""
    import torch
    from noahmp_b200.engine import NoahMP
    w = synth.make_workload(name, nsteps=nsteps)
    order = divergence_order(w.static_i, w.static_f[L.C["LAT"]], 1) if w.ncol > 1 else np.zeros(1, np.int64)
    state = np.ascontiguousarray(w.state[:, order])
    eng = NoahMP(ParamTables.default(), w.opts, device=dev.index, fast=variant)
    eng.set_stream(stream)
    with torch.cuda.stream(stream):
        st = torch.from_numpy(state).to(dev)
        sf = torch.from_numpy(np.ascontiguousarray(w.static_f[:, order])).to(dev)
        si = torch.from_numpy(np.ascontiguousarray(w.static_i[order])).to(dev)
        sincos = torch.from_numpy(np.ascontiguousarray(w.sincos[:, order])).to(dev)
        ids = torch.from_numpy(order.astype(np.int64)).to(dev)
        chunk = 24
        fo = torch.empty((chunk, L.NFORCING, w.ncol), dtype=torch.float32, device=dev)
        bd = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    coef = synth.synth_params_array(w.synth)
    evs = []
    l0 = eng.launch_count()
    # the generator hashes each column's GENERATOR id (ids), so the re-ordered columns keep their own forcing
    with torch.cuda.stream(stream):
        for s0 in range(0, w.nsteps, chunk):
            n = min(chunk, w.nsteps - s0)
            for s in range(n):
                sc = synth.step_scalars(w, s0 + s)
                eng.synth_forcing_ids(ids, w.synth.seed, s0 + s, coef,
                                      np.asarray([sc["sd"], sc["cdch"], sc["t_diur"], sc["t_seas"]], np.float32), sincos, fo[s])
            rp = eng.run_params(w.dt, synth.ZSOIL, w.julian0 + s0 * w.dt / 86400.0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.run(rp, st, fo[:n], sf, si, None, bd)
            e1.record(stream)
            evs.append((e0, e1))
    stream.synchronize()
    kern_s = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    cs = w.ncol * w.nsteps
    b = bd.cpu().numpy()
    out = {"workload": name, "columns": w.ncol, "model_steps": w.nsteps, "dt_s": w.dt, "cell_steps": cs, "variant": variant,
           "value": cs / kern_s, "unit": UNIT, "kernel_s": kern_s, "launches": int(eng.launch_count() - l0 - w.nsteps),
           "roofline_frac": B_ALG * cs / kern_s / 1e9 / peak, "flagged_column_launches": int(b[L.B["NFLAGGED"]]),
           "max_abs_water_error_mm": float(b[L.B["ERRWAT_MAXABS"]]), "sneqv_mean_mm": float(b[L.B["SNEQV"]] / max(b[L.B["NCELL"]], 1.0)),
           "timing": "CUDA events around the step launches only (forcing generation is outside)"}
    iss = issue_roofline(cs / kern_s, variant, sm_count, clocks_mhz)
    if iss:
        out["issue_frac_at_conus_instruction_count"] = iss["frac"]
    if twin_name:
        ts = twin_sample(twin_name, order, 0, w.ncol)
        if ts is not None and w.nsteps == int(ts[0]["nsteps"]) and w.ncol == int(ts[0]["ncol"]):
            g, sel, pos = ts
            out["parity"] = parity_of(st[:, torch.from_numpy(pos).to(dev)].cpu().numpy(), g, sel)
    eng.close()
    del st, sf, si, fo
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from noahmp_b200.engine import NoahMP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU (no CPU fallback exists)")
    local = device_for_rank(local_rank, world)       # the GPU of this rank (ranks are spread over the node's host bridges)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = build_workload(args)
    order = w.order
    c0, nloc = shard(w.ncol, rank, world)
    ws = slice_workload(w, c0, nloc)
    eng = NoahMP(ParamTables.default(), w.opts, device=local, fast=args.variant)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream)
    rp_of = lambda j0: eng.run_params(w.dt, synth.ZSOIL, j0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)  # noqa: E731

    # ---- inputs resident in HBM: state, static, forcing of every model step (generated on device) -----
    with torch.cuda.stream(stream):
        state0 = torch.from_numpy(ws.state).to(dev)
        state = state0.clone()
        static_f = torch.from_numpy(ws.static_f).to(dev)
        static_i = torch.from_numpy(ws.static_i).to(dev)
        sincos = torch.from_numpy(ws.sincos).to(dev)
        ids = torch.from_numpy(ws.order).to(dev)
        forcing = torch.empty((w.nsteps, L.NFORCING, nloc), dtype=torch.float32, device=dev)
        coef = synth.synth_params_array(w.synth)
        scal = []
        for s in range(w.nsteps):
            sc = synth.step_scalars(w, s)
            scal.append(sc)
            eng.synth_forcing_ids(ids, w.synth.seed, s, coef,
                                  np.asarray([sc["sd"], sc["cdch"], sc["t_diur"], sc["t_seas"]], np.float32), sincos, forcing[s])
        budget = torch.zeros(L.NBUDGET, dtype=torch.float64, device=dev)
    stream.synchronize()

    def one_step(k):
        # the simulation simply continues: bench step k starts model time at julian0 + k*model_steps*dt
        eng.run(rp_of(w.julian0 + k * w.nsteps * w.dt / 86400.0), state, forcing, static_f, static_i)

    with torch.cuda.stream(stream):
        for k in range(args.warmup):
            one_step(k)
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = eng.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        evs[0].record(stream)
        for k in range(args.steps):
            one_step(args.warmup + k)
            evs[k + 1].record(stream)
    stream.synchronize()
    torch.cuda.synchronize()
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    t_max = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_max.item())
    peak, peak_src = measured_peak_gbs()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    # ---- auxiliary kernel of the widened path: forcing preparation, pure HBM streaming
    aux = {}
    try:
        from noahmp_b200 import forcing as FRC
        geo_p = torch.from_numpy(FRC.geo_planes(ws.static_f[L.C["LAT"]], np.zeros(nloc))).to(dev)
        out_f = torch.empty((L.NFORCING, nloc), dtype=torch.float32, device=dev)
        r0 = forcing[0, :8].contiguous()
        r1 = forcing[min(1, w.nsteps - 1), :8].contiguous()
        sc = FRC.solar_scalars(__import__("datetime").datetime(2000, 7, 1, 18))
        with torch.cuda.stream(stream):
            for _ in range(3):
                eng.prepare_forcing(r0, r1, 0.25, geo_p, sc, out_f)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            for _ in range(10):
                eng.prepare_forcing(r0, r1, 0.25, geo_p, sc, out_f)
            p1.record(stream)
        stream.synchronize()
        pms = p0.elapsed_time(p1) / 10.0
        pbytes = 4 * nloc * (16 + 4 + 9)
        aux["prepare_forcing"] = {"ms": pms, "bytes_per_launch": pbytes, "achieved_gbs": pbytes / (pms * 1e-3) / 1e9,
                                  "frac_of_measured_hbm": pbytes / (pms * 1e-3) / 1e9 / peak, "bound": "hbm",
                                  "note": "29 words per column, 128-bit accesses; buffers (2 GB at 17e6 columns) exceed L2; the "
                                          "records-driven launches of `e2e` do this arithmetic inside the step kernel instead"}
        del geo_p, out_f, r0, r1
    except Exception as exc:  # pragma: no cover - diagnostics only
        aux["prepare_forcing"] = {"error": str(exc)}

    # ---- the other two builds of the same source on the same resident inputs (transparency: `value` is the fast
    # build; the IEEE-arithmetic build and the fp64-libm verification build are reported next to it)
    other = {}
    if not args.no_other_variants:
        for v in ("parity", "strict"):
            if v == args.variant:
                continue
            eng_v = NoahMP(ParamTables.default(), w.opts, device=local, fast=v)
            eng_v.set_stream(stream)
            st_v = state0.clone()
            rp0 = rp_of(w.julian0)
            with torch.cuda.stream(stream):
                eng_v.run(rp0, st_v, forcing, static_f, static_i)      # warm-up
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for k in range(2):
                    eng_v.run(rp0, st_v, forcing, static_f, static_i)
                e1.record(stream)
            stream.synchronize()
            tv = torch.tensor([e0.elapsed_time(e1) / 2.0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            other[v] = {"value": w.ncol * w.nsteps / (float(tv.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(tv.item()),
                        "kernel": eng_v.kernel_info()}
            eng_v.close()
            del st_v

    # ---- one extra pass from the initial state with the budget sums: domain diagnostics, and the state every parity figure
    # of this line is computed on (24 model steps from the generator's initial columns)
    with torch.cuda.stream(stream):
        chk = state0.clone()
        eng.run(rp_of(w.julian0), chk, forcing, static_f, static_i, None, budget)
    stream.synchronize()
    bud = budget_means(allreduce_budget(budget.clone()).cpu().numpy(), w.nsteps)
    parity = None
    try:
        if w.name == "conus1km" and args.ncol is None and args.model_steps is None:
            ts = twin_sample("conus1km_day", order, c0, nloc)
            if ts is not None and len(ts[1]):
                g, sel, pos = ts
                parity = parity_of(chk[:, torch.from_numpy(pos).to(dev)].cpu().numpy(), g, sel)
                parity["rank0_columns_of_sample"] = int(len(sel))
    except Exception as exc:  # pragma: no cover - the line must survive a missing fixture
        parity = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- e2e through the host-facing C-ABI entry points -------------------------------------------------
    # (1) headline `e2e`: noahmp_advance_records_host -- the time-marching pattern: column state, static and position planes
    #     resident on the device (uploaded once), every model step's forcing RECORD (8 fp32 planes) copied from PINNED host
    #     memory inside the timed region (ring of device slots, copies overlapped with compute), time interpolation and COSZ
    #     done in the step kernel's load, the 16 fp64 budget sums read back and all-reduced over the ranks every bench step.
    # (2) `e2e_state_roundtrip`: noahmp_run_host -- additionally moves the whole state H2D and D2H every step.
    e2e = None
    e2e_rt = None
    e2e_nml = None
    probe = None
    if not args.no_e2e:
        probe = h2d_probe(dev, world)
        nplanes = L.NSTATE_CORE if w.opts == synth.DEFAULT_OPTS else L.NSTATE
        h_records = torch.empty((w.nsteps, 8, nloc), dtype=torch.float32).pin_memory()
        h_records.copy_(forcing[:, :8])
        h_budget = torch.zeros(L.NBUDGET, dtype=torch.float64).pin_memory()
        # position planes of the synthetic grid: sin(lat), cos(lat) as the generator has them, longitude 0; the generator's COSZ
        # = sin(lat) sd + cos(lat) (cos(decl) cos(H)) is the kernel's formula with cd = 1, ch0 = cos(decl) cos(H), sh0 = 0
        geo = torch.stack([sincos[0], sincos[1], torch.zeros_like(sincos[0]), torch.ones_like(sincos[0])]).contiguous()
        steps_tab = eng.step_table(list(range(w.nsteps)), list(range(w.nsteps)), [0.0] * w.nsteps,
                                   [(sc["sd"], 1.0, sc["cdch"], 0.0) for sc in scal])
        torch.cuda.synchronize()
        ne2e = max(1, min(args.steps, 3))
        spl = args.e2e_steps_per_launch

        # host threads that zero-suppress the sparse record planes (noahmp_set_record_packing): this rank's share of the cores
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        pack_threads = args.pack_threads or max(1, min(16, ncpu // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))

        def e2e_leg(st0, recs_host, geo_b, sf_b, si_b, resident_end, pack_mask):
            """ne2e bench steps through noahmp_advance_records_host on one block of columns; returns (seconds, max over ranks),
            launches, flagged, whether the warm-up step reproduced the resident path bit for bit, and the bytes that crossed
            the host link per bench step (all ranks)."""
            eng.set_record_packing(pack_mask, pack_threads)
            st_e = st0.clone()
            eng.advance_records_host(rp_of(w.julian0), st_e, recs_host, steps_tab, geo_b, sf_b, si_b, h_budget, spl)  # warm-up
            same_ = bool(torch.equal(st_e[:nplanes].view(torch.int32), resident_end[:nplanes].view(torch.int32)))
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            l0_ = eng.launch_count()
            b0_ = eng.h2d_bytes()
            p0_ = eng.pack_seconds()
            t0_ = time.perf_counter()
            flagged_ = 0
            gsum_ = None
            for k in range(ne2e):
                flagged_ += eng.advance_records_host(rp_of(w.julian0 + (k + 1) * w.nsteps * w.dt / 86400.0), st_e, recs_host, steps_tab,
                                                     geo_b, sf_b, si_b, h_budget, spl)
                # the path's only collective, at diagnostics cadence (one model day), inside the timed region
                gsum_ = allreduce_budget(h_budget.to(dev, non_blocking=True))
            if gsum_ is not None:
                gsum_ = gsum_.cpu()
            t_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device=dev)
            nb_ = torch.tensor([float(eng.h2d_bytes() - b0_) / ne2e], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                dist.all_reduce(nb_, op=dist.ReduceOp.SUM)
            e2e_leg.pack_s = (eng.pack_seconds() - p0_) / ne2e     # this rank's host time spent packing, per bench step
            return float(t_.item()), int(eng.launch_count() - l0_), int(flagged_), same_, int(nb_.item())

        # The records cross the host link either as they are (eight dense planes) or with the planes that are zero over much of
        # the grid (night-time SOLDN, dry PRCP) zero-suppressed by host threads - lossless, the kernel's load reads the packed
        # form.  Both legs run; the faster one is the headline (packing costs host memory bandwidth, which a box may not have
        # to spare when many ranks share it), the other stays beside it.
        PACK = 0xA0
        dt_d, launches_d, flagged_d, same_d, bytes_d = e2e_leg(state0, h_records, geo, static_f, static_i, chk, 0)
        dt_p, launches_p, flagged_p, same_p, bytes_p = e2e_leg(state0, h_records, geo, static_f, static_i, chk, PACK)
        pack_s = e2e_leg.pack_s
        sweep = {}
        for th_ in [int(x) for x in args.pack_sweep.split(",") if x.strip()]:
            keep = pack_threads
            pack_threads = th_
            r_ = e2e_leg(state0, h_records, geo, static_f, static_i, chk, PACK)
            sweep[str(th_)] = {"value": w.ncol * w.nsteps * ne2e / r_[0], "host_pack_seconds_per_step": e2e_leg.pack_s}
            pack_threads = keep
        packed_wins = dt_p < dt_d
        pack_mask = PACK if packed_wins else 0
        dt_e, launches_e, flagged, same, bytes_e = ((dt_p, launches_p, flagged_p, same_p, bytes_p) if packed_wins else
                                                    (dt_d, launches_d, flagged_d, same_d, bytes_d))
        tpk = time.perf_counter()
        from noahmp_b200.engine import pack_plane
        nnz_pk = pack_plane(h_records[0, 7].numpy(), pack_threads)[2].shape[0]
        tpk = time.perf_counter() - tpk
        e2e = {"value": w.ncol * w.nsteps * ne2e / dt_e, "unit": UNIT,
               "h2d_bytes_per_step": bytes_e, "d2h_bytes_per_step": 8 * L.NBUDGET * world,
               "steps": ne2e, "api": "noahmp_advance_records_host", "steps_per_launch": spl,
               "gpu_launches": launches_e, "flagged": flagged, "partition": "equal blocks",
               "state_equals_resident_path": same, "collective": "budget all-reduce (15 x SUM + 1 x MAX, fp64) once per bench step, timed",
               "records": "SOLDN and PRCP planes zero-suppressed on the host (lossless)" if packed_wins else "eight dense planes",
               "dense_records": {"value": w.ncol * w.nsteps * ne2e / dt_d, "h2d_bytes_per_step": bytes_d,
                                 "state_equals_resident_path": same_d},
               "packed_records": {"value": w.ncol * w.nsteps * ne2e / dt_p, "h2d_bytes_per_step": bytes_p,
                                  "state_equals_resident_path": same_p, "host_threads_per_rank": pack_threads,
                                  "host_pack_seconds_per_step": pack_s, "thread_sweep": sweep or None,
                                  "one_plane_alone": {"columns": nloc, "nonzero": int(nnz_pk), "seconds": tpk,
                                                      "note": "PRCP plane of the first record through noahmp_pack_plane (allocations included)"}},
               "host_link": probe, "cpu_binding": numa,
               "note": "state resident on device; forcing RECORDS (8 planes) streamed from pinned host memory, each once; time "
                       "interpolation + COSZ in the step kernel's load; budget read back and all-reduced; h2d_bytes_per_step is "
                       "counted by the library from the copies it issues"}
        # Where the ranks' host links differ (measured by the probe while all ranks copy), a run that streams its forcing is as
        # fast as its slowest link allows: the same leg again with the columns dealt out in proportion to each rank's link.  It
        # becomes the headline `e2e` (the equal-block figure stays beside it); the device-timed `value` keeps equal blocks.
        rates = probe.get("h2d_gbs_each_rank_while_all_copy") or []
        if world > 1 and rates and max(rates) > 1.15 * min(rates):
            blk, err = None, None
            try:     # preparation may fail on one rank (memory): the ranks agree below before any of them enters the leg
                wsum = sum(rates)
                edges = [0]
                for r in range(world - 1):
                    edges.append(min(w.ncol, int(round(w.ncol * sum(rates[:r + 1]) / wsum / 256.0)) * 256))
                edges.append(w.ncol)
                c0w, nw = edges[rank], edges[rank + 1] - edges[rank]
                if nw <= 0:
                    raise ValueError("empty block")
                wb = slice_workload(w, c0w, nw)
                with torch.cuda.stream(stream):
                    st_w = torch.from_numpy(wb.state).to(dev)
                    sf_w = torch.from_numpy(wb.static_f).to(dev)
                    si_w = torch.from_numpy(wb.static_i).to(dev)
                    sc_w = torch.from_numpy(wb.sincos).to(dev)
                    ids_w = torch.from_numpy(wb.order).to(dev)
                    fo_w = torch.empty((w.nsteps, L.NFORCING, nw), dtype=torch.float32, device=dev)
                    for s_ in range(w.nsteps):
                        sc = scal[s_]
                        eng.synth_forcing_ids(ids_w, w.synth.seed, s_, coef,
                                              np.asarray([sc["sd"], sc["cdch"], sc["t_diur"], sc["t_seas"]], np.float32), sc_w, fo_w[s_])
                    end_w = st_w.clone()
                    eng.run(rp_of(w.julian0), end_w, fo_w, sf_w, si_w)          # the resident path on this block (the check)
                stream.synchronize()
                rec_w = torch.empty((w.nsteps, 8, nw), dtype=torch.float32).pin_memory()
                rec_w.copy_(fo_w[:, :8])
                geo_w = torch.stack([sc_w[0], sc_w[1], torch.zeros_like(sc_w[0]), torch.ones_like(sc_w[0])]).contiguous()
                del fo_w
                torch.cuda.synchronize()
                blk = (st_w, rec_w, geo_w, sf_w, si_w, end_w)
            except Exception as exc:  # pragma: no cover - the equal-block figure stands
                err = f"{type(exc).__name__}: {exc}"
            ready = torch.tensor([1 if blk is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ready, op=dist.ReduceOp.MIN)
            if int(ready.item()) == 1:
                dt_w, launches_w, flagged_w, same_w, bytes_w = e2e_leg(*blk, pack_mask)
                equal = {k: e2e[k] for k in ("value", "gpu_launches", "flagged", "state_equals_resident_path", "h2d_bytes_per_step")}
                e2e.update({"value": w.ncol * w.nsteps * ne2e / dt_w, "gpu_launches": launches_w, "flagged": flagged_w,
                            "h2d_bytes_per_step": bytes_w,
                            "state_equals_resident_path": bool(same_w and same),
                            "partition": "blocks in proportion to each rank's measured host link",
                            "columns_per_rank": [edges[r + 1] - edges[r] for r in range(world)], "equal_blocks": equal})
            else:
                e2e["link_weighted_error"] = err or "another rank could not prepare its block"
            del blk
        # (1b) the same entry point at the cadence of the reference's own namelist (run/case.nml:27 interval_seconds = 900,
        #      :7 input_frequency = '1 hour'): four model steps per record, interpolated in the kernel's load, so a record that
        #      crossed the link serves four steps - 8 B (or less, packed) per cell-step instead of 32.  Extra figure, not the headline.
        try:
            import copy
            spr = 4
            w900 = copy.copy(w)
            w900.dt = w.dt / spr
            n900 = w.nsteps * spr
            sc900 = [synth.step_scalars(w900, s_) for s_ in range(n900)]
            tab900 = eng.step_table([s_ // spr for s_ in range(n900)], [(s_ // spr + 1) % w.nsteps for s_ in range(n900)],
                                    [(s_ % spr) / spr for s_ in range(n900)], [(sc["sd"], 1.0, sc["cdch"], 0.0) for sc in sc900])
            rp900 = lambda j0: eng.run_params(w900.dt, synth.ZSOIL, j0, w.yearlen, w.zref, w.co2_frac, w.o2_frac, w.foln)  # noqa: E731
            eng.set_record_packing(pack_mask, pack_threads)
            st_n = state0.clone()
            eng.advance_records_host(rp900(w.julian0), st_n, h_records, tab900, geo, static_f, static_i, h_budget, spr)   # warm-up
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            b0n, l0n, t0n = eng.h2d_bytes(), eng.launch_count(), time.perf_counter()
            flagged_n = eng.advance_records_host(rp900(w.julian0 + 1.0), st_n, h_records, tab900, geo, static_f, static_i, h_budget, spr)
            allreduce_budget(h_budget.to(dev, non_blocking=True)).cpu()
            t_n = torch.tensor([time.perf_counter() - t0n], dtype=torch.float64, device=dev)
            nb_n = torch.tensor([float(eng.h2d_bytes() - b0n)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_n, op=dist.ReduceOp.MAX)
                dist.all_reduce(nb_n, op=dist.ReduceOp.SUM)
            e2e_nml = {"value": w.ncol * n900 / float(t_n.item()), "unit": UNIT, "dt_s": w900.dt, "record_interval_s": w.dt,
                       "model_steps": n900, "steps_per_launch": spr, "h2d_bytes_per_step": int(nb_n.item()),
                       "gpu_launches": int(eng.launch_count() - l0n), "flagged": int(flagged_n), "api": "noahmp_advance_records_host",
                       "note": "one model day at the reference namelist's cadence (900 s steps, hourly records interpolated in the "
                               "kernel's load), same records and entry point as `e2e`"}
            del st_n
        except Exception as exc:  # pragma: no cover - an extra figure must not cost the line
            e2e_nml = {"error": f"{type(exc).__name__}: {exc}"}
        # (2) full state round trip
        h_forcing = torch.empty((w.nsteps, L.NFORCING, nloc), dtype=torch.float32).pin_memory()
        h_forcing.copy_(forcing)
        h_state = torch.from_numpy(ws.state.copy()).pin_memory()
        h_sf = torch.from_numpy(ws.static_f).pin_memory()
        h_si = torch.from_numpy(ws.static_i).pin_memory()
        eng.run_host(rp_of(w.julian0), h_state, h_forcing, h_sf, h_si)       # warm-up (allocates device staging)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(ne2e):
            eng.run_host(rp_of(w.julian0 + (k + 1) * w.nsteps * w.dt / 86400.0), h_state, h_forcing, h_sf, h_si)
        dt_r = time.perf_counter() - t0
        t_r = torch.tensor([dt_r], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_r, op=dist.ReduceOp.MAX)
        e2e_rt = {"value": w.ncol * w.nsteps * ne2e / float(t_r.item()), "unit": UNIT,
                  "h2d_bytes_per_step": 4 * w.ncol * (nplanes + w.nsteps * L.NFORCING + L.NSTATIC_F + 1),
                  "d2h_bytes_per_step": 4 * w.ncol * nplanes, "steps": ne2e, "api": "noahmp_run_host"}
        del h_state, h_forcing, h_records

    # ---- the other BASELINE configs at full size and length on this GPU (N = 1 only: they fit one GPU many times over)
    others = None
    if world == 1 and not args.no_other_workloads and w.name == "conus1km" and args.ncol is None:
        del forcing, state, chk, state0
        torch.cuda.empty_cache()
        others = {}
        mhz = (clocks or {}).get("sm_mhz")
        for name, nsteps, twin in (("snow1m", 720, "snow1m_month"), ("global025", 8760, "global025_year"),
                                   ("uniform100k", 48, "uniform100k"), ("single", 96, None)):
            try:
                others[name] = run_other_workload(name, nsteps, twin, args.variant, dev, stream, peak, sm_count, mhz)
            except Exception as exc:  # pragma: no cover
                others[name] = {"error": f"{type(exc).__name__}: {exc}"}
        try:   # the run-time-options kernel once (all 12 options as kernel arguments, 57-plane state): opt_tbot = 2 costs the same physics
            wopt = synth.make_workload("conus1km", ncol=2_000_000, nsteps=24, opts=[1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 1])
            o2 = divergence_order(wopt.static_i, wopt.static_f[L.C["LAT"]], 1)
            eng_o = NoahMP(ParamTables.default(), wopt.opts, device=local, fast=args.variant)
            eng_o.set_stream(stream)
            with torch.cuda.stream(stream):
                st_o = torch.from_numpy(np.ascontiguousarray(wopt.state[:, o2])).to(dev)
                sf_o = torch.from_numpy(np.ascontiguousarray(wopt.static_f[:, o2])).to(dev)
                si_o = torch.from_numpy(np.ascontiguousarray(wopt.static_i[o2])).to(dev)
                fo_o = torch.from_numpy(np.ascontiguousarray(synth.forcing_all(wopt)[:, :, o2])).to(dev)
                rp_o = eng_o.run_params(wopt.dt, synth.ZSOIL, wopt.julian0, wopt.yearlen, wopt.zref, wopt.co2_frac, wopt.o2_frac, wopt.foln)
                eng_o.run(rp_o, st_o, fo_o, sf_o, si_o)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(3):
                    eng_o.run(rp_o, st_o, fo_o, sf_o, si_o)
                e1.record(stream)
            stream.synchronize()
            others["runtime_options_kernel"] = {"workload": "conus1km", "columns": wopt.ncol, "model_steps": 24, "options": wopt.opts,
                                                "value": wopt.ncol * 24 * 3 / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
                                                "kernel": eng_o.kernel_info(), "note": "OptRuntime instantiation, 57-plane state"}
            eng_o.close()
        except Exception as exc:  # pragma: no cover
            others["runtime_options_kernel"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        value = w.ncol * w.nsteps * args.steps / (total_ms_max * 1e-3)
        k_ms = float(np.mean(per_launch_ms))
        achieved = B_ALG * nloc * w.nsteps / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(nloc, w.nsteps, args.variant)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": config_of(w, args, {
                    "columns_per_gpu": nloc, "variant": args.variant, "partition": f"block x{world}",
                    "gpus": [device_for_rank(r, world) for r in range(world)],
                    "l2": "inputs larger than L2 (state+forcing per launch >> 126 MB); no flush needed"
                          if nloc * 4 * (2 * 44 + 9 * w.nsteps) > 2 * 126e6 else "inputs fit L2: traffic is L2-resident",
                    "kernel": eng.kernel_info()}),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write)",
                             "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": B_ALG * nloc * w.nsteps, "peak_source": peak_src, "bytes_per_cell_step": B_ALG,
                             "kernel": "column_step_kernel", "kernel_ms": k_ms,
                             "note": "kernel is instruction-issue bound (see profiles/); fraction is reported against HBM as "
                                     "the contract asks"},
                "e2e": e2e, "e2e_namelist_cadence": e2e_nml, "e2e_state_roundtrip": e2e_rt, "parity": parity, "other_variants": other, "other_workloads": others,
                "aux_kernels": aux, "gpu_launches": int(launches), "clocks": clocks, "budget": bud}
        line["roofline_issue"] = issue_roofline(nloc * w.nsteps / (k_ms * 1e-3), args.variant, sm_count, (clocks or {}).get("sm_mhz"))
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(w, args.cpu_sample_cols)
            tr = translated_reference_leg(w)
            if tr:
                line["cpu_baseline"]["translated_reference"] = tr
            if w.name == "conus1km" and args.ncol is None and not args.no_other_workloads:
                per = {}
                for name in ("single", "uniform100k", "global025", "conus1km", "snow1m"):
                    try:
                        per[name] = cpu_leg_of(name, args.variant, local)
                    except Exception as exc:  # pragma: no cover - a CPU leg must never take the line down
                        per[name] = {"error": f"{type(exc).__name__}: {exc}"}
                line["cpu_baseline"]["per_config"] = per
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else this process prints (NCCL banner, torchrun
    notices, library chatter) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
