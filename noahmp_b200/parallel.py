"""Multi-GPU layout of the column step: contiguous block partition, no halo, one tiny reduction.

Columns share nothing but read-only tables (``noahmp_sflx`` is a one-column routine, func:66; ILOC/JLOC only
feed messages), so each rank owns ``ceil(ncol/world)`` consecutive columns and the data path has NO
collective.  The only exchange is the domain-mean water/energy budget: 15 fp64 sums + 1 fp64 max
(``layout.BUDGET``), all-reduced over NCCL on GPUs (gloo in the CPU tests) at diagnostics cadence.
"""
from __future__ import annotations

from . import layout as L


def shard(ncol: int, rank: int, world: int) -> tuple[int, int]:
    """(first column, column count) of ``rank``; trailing ranks may get fewer (or zero) columns."""
    per = (ncol + world - 1) // world
    c0 = min(rank * per, ncol)
    return c0, min(c0 + per, ncol) - c0


def device_for_rank(local_rank: int, world: int, ndev: int | None = None) -> int:
    """GPU of this node for a local rank.  With as many visible GPUs as ranks: the rank itself.  With more GPUs than ranks
    the ranks are spread over the node with a constant stride (4 ranks on 8 GPUs: 0, 2, 4, 6), so that they do not all sit
    behind the same host bridge: on the 8 x B200 box measured (`profiles/r02_logs/r02_h2d_topology.jsonl`) four ranks copying
    from pinned host memory get 115 GB/s together on GPUs 0-3 and 213 GB/s on GPUs 0, 2, 4, 6 - the step path has no
    inter-GPU traffic, so the host link is the only placement concern.  ``NOAHMP_GPU_MAP=4,5,6,7`` overrides."""
    import os
    explicit = os.environ.get("NOAHMP_GPU_MAP")
    if explicit:
        m = [int(x) for x in explicit.split(",")]
        return m[local_rank % len(m)]
    if ndev is None:
        import torch
        ndev = torch.cuda.device_count()
    if world <= 1 or ndev < 2 * world:
        return local_rank
    return local_rank * (ndev // world)


def divergence_order(static_i, lat, world: int = 1, chunk: int = 32768):
    """Permutation of the columns that (a) groups them by divergence class -- land-use type (canopy tile on/off,
    table rows), then latitude band (day/night, phenology month) -- so the 32 lanes of a warp take the same
    branches, and (b) deals the sorted sequence to ``world`` ranks in round-robin chunks, so the contiguous block
    partition of :func:`shard` gives every rank the same class mix (cheap barren columns are not all on one GPU).
    Columns are independent (func:66 is a one-column routine), so their order is the driver's choice; apply the
    permutation once to state/static planes at set-up and invert it on output."""
    import numpy as np
    static_i = np.asarray(static_i)
    lut = (static_i & 0xff).astype(np.int64)
    band = np.clip(((np.asarray(lat, np.float64) + 1.6) * 20.0).astype(np.int64), 0, 63)
    order = np.argsort(lut * 64 + band, kind="stable")
    if world <= 1:
        return order
    n = order.size
    nchunk = (n + chunk - 1) // chunk
    owner = np.arange(nchunk) % world
    pieces = [[] for _ in range(world)]
    for k in range(nchunk):
        pieces[owner[k]].append(order[k * chunk:(k + 1) * chunk])
    return np.concatenate([np.concatenate(p) if p else np.empty(0, order.dtype) for p in pieces])


def allreduce_budget(budget, group=None):
    """In-place all-reduce of a [NBUDGET] fp64 tensor: SUM for every slot except ERRWAT_MAXABS (MAX)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return budget
    imax = L.B["ERRWAT_MAXABS"]
    mx = budget[imax].clone()
    budget[imax] = 0.0
    dist.all_reduce(budget, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    budget[imax] = mx
    return budget


def gather_columns(local, ncol: int, group=None):
    """[P][nloc] shards of the block partition of :func:`shard` -> the full [P][ncol] tensor on every rank (all-gather of
    equal-size padded blocks; used at output / restart cadence only, never on the step path)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (ncol + world - 1) // world
    if local.shape[1] == per:
        pad = local.contiguous()
    else:
        pad = torch.zeros((local.shape[0], per), dtype=local.dtype, device=local.device)
        pad[:, :local.shape[1]] = local
    out = torch.empty((world * pad.shape[0], per), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)     # one collective into one buffer: rank-major [world * P][per]
    return out.view(world, pad.shape[0], per).permute(1, 0, 2).reshape(local.shape[0], world * per)[:, :ncol].contiguous()


def budget_means(budget, nsteps: int) -> dict:
    """Domain means per cell (and per step for fluxes) from the reduced sums."""
    b = {n: float(budget[i]) for i, n in enumerate(L.BUDGET)}
    n = max(b["NCELL"], 1.0)
    out = {"cells": int(b["NCELL"]), "flagged": int(b["NFLAGGED"]), "errwat_mean_mm": b["ERRWAT_SUM"] / n / max(nsteps, 1),
           "errwat_maxabs_mm": b["ERRWAT_MAXABS"]}
    for k in ("FSA", "FSR", "FIRA", "FSH", "LH", "SSOIL"):
        out[k.lower() + "_wm2"] = b[k] / n / max(nsteps, 1)
    for k in ("PRCP_DT", "ET_DT", "RUNOFF_DT"):
        out[k.lower().replace("_dt", "_mm")] = b[k] / n
    out["sneqv_mm"] = b["SNEQV"] / n
    out["soilwater_mm"] = b["SOILWATER"] / n
    return out
