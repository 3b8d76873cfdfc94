"""N>1 host logic on CPU (gloo, world_size 2): block partition, shard-independent forcing, and the one
collective of the path (budget all-reduce: SUM of 15 slots + MAX of one)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from noahmp_b200 import layout as L, synth
from noahmp_b200.parallel import allreduce_budget, budget_means, divergence_order, gather_columns, shard


def test_shard_covers_everything_once():
    for ncol in (1, 7, 100, 17_000_001):
        for world in (1, 2, 3, 8):
            parts = [shard(ncol, r, world) for r in range(world)]
            assert sum(n for _, n in parts) == ncol
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(n for _, n in parts) - min(n for _, n in parts if n or world <= ncol) <= (ncol + world - 1) // world


def test_divergence_order_is_a_balanced_permutation():
    w = synth.make_workload("conus1km", ncol=200_000, nsteps=1)
    lat = w.static_f[L.C["LAT"]]
    for world in (1, 2, 8):
        order = divergence_order(w.static_i, lat, world, chunk=512)
        assert np.array_equal(np.sort(order), np.arange(w.ncol))
        lut = (w.static_i & 0xff)[order]
        if world == 1:
            assert (np.diff(lut) >= 0).all()                       # grouped by land-use type
        shares = []
        for r in range(world):
            c0, n = shard(w.ncol, r, world)
            shares.append(np.bincount(lut[c0:c0 + n], minlength=28) / n)
        shares = np.asarray(shares)
        assert np.abs(shares - shares.mean(axis=0)).max() < 0.03     # every rank sees the same class mix
        runs = 1 + int((np.diff(lut) != 0).sum())
        assert runs <= 28 * (w.ncol // 512 + world + 1)               # still long homogeneous runs inside a rank


def _budget_from_oracle(w, c0, n, forcing):
    """Run the CPU oracle on one shard, step by step, and accumulate the same 16 budget slots the kernel does."""
    import sys
    sys.path.insert(0, ROOT)
    from noahmp_b200.tables import ParamTables
    from oracle.oracle import Oracle
    o = Oracle(ParamTables.default(), w.opts)
    st = np.ascontiguousarray(w.state[:, c0:c0 + n])
    sf = np.ascontiguousarray(w.static_f[:, c0:c0 + n])
    si = np.ascontiguousarray(w.static_i[c0:c0 + n])
    b = np.zeros(L.NBUDGET)
    diag = np.zeros((L.NDIAG, n), np.float32)
    flagged = np.zeros(n, bool)
    for s in range(w.nsteps):
        fs = np.ascontiguousarray(forcing[s:s + 1, :, c0:c0 + n])
        o.run(st, fs, sf, si, dt=w.dt, zsoil=synth.ZSOIL, julian0=w.julian0 + s * w.dt / 86400.0, yearlen=w.yearlen, diag=diag)
        d = lambda k: diag[L.D[k]].astype(np.float64)  # noqa: E731
        b[L.B["ERRWAT_SUM"]] += d("ERRWAT").sum()
        b[L.B["ERRWAT_ABS"]] += np.abs(d("ERRWAT")).sum()
        b[L.B["ERRWAT_MAXABS"]] = max(b[L.B["ERRWAT_MAXABS"]], np.abs(d("ERRWAT")).max())
        for k in ("FSA", "FSR", "FIRA", "FSH", "SSOIL"):
            b[L.B[k]] += d(k).sum()
        b[L.B["LH"]] += (d("FCEV") + d("FGEV") + d("FCTR")).sum()
        b[L.B["PRCP_DT"]] += (fs[0, L.F["PRCP"]].astype(np.float64) * w.dt).sum()
        flagged |= diag[L.D["FLAGS"]].view(np.int32) != 0
    b[L.B["NCELL"]] = n
    b[L.B["NFLAGGED"]] = flagged.sum()
    b[L.B["SNEQV"]] = st[L.S["SNEQV"]].astype(np.float64).sum()
    return b


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workload("global025", ncol=301, nsteps=6)
    c0, n = shard(w.ncol, rank, world)
    import copy
    ws = copy.copy(w)
    ws.sincos = np.ascontiguousarray(w.sincos[:, c0:c0 + n])
    forcing_local = np.stack([synth.forcing_step(ws, s, col0=c0) for s in range(w.nsteps)])
    full = np.zeros((w.nsteps, L.NFORCING, w.ncol), np.float32)
    full[:, :, c0:c0 + n] = forcing_local
    b = torch.from_numpy(_budget_from_oracle(w, c0, n, full))
    allreduce_budget(b)
    if rank == 0:
        q.put(b.numpy().copy())
    dist.destroy_process_group()


def test_budget_allreduce_matches_single_process(oracle_libs):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.make_workload("global025", ncol=301, nsteps=6)
    ref = _budget_from_oracle(w, 0, w.ncol, synth.forcing_all(w))
    assert got[L.B["NCELL"]] == 301 and got[L.B["NFLAGGED"]] == ref[L.B["NFLAGGED"]] == 0
    assert got[L.B["ERRWAT_MAXABS"]] == ref[L.B["ERRWAT_MAXABS"]]          # MAX, not SUM
    for k in ("FSA", "FSR", "FIRA", "FSH", "LH", "SSOIL", "PRCP_DT", "SNEQV", "ERRWAT_SUM", "ERRWAT_ABS"):
        assert np.isclose(got[L.B[k]], ref[L.B[k]], rtol=1e-12, atol=1e-9), k
    m = budget_means(got, w.nsteps)
    assert m["cells"] == 301 and abs(m["errwat_mean_mm"]) < 1e-3


def test_allreduce_is_identity_without_process_group():
    b = torch.arange(L.NBUDGET, dtype=torch.float64)
    assert torch.equal(allreduce_budget(b.clone()), b)


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ncol = 301                                                  # uneven: 151 + 150
    full = torch.arange(5 * ncol, dtype=torch.float32).reshape(5, ncol)
    c0, n = shard(ncol, rank, world)
    got = gather_columns(full[:, c0:c0 + n].contiguous(), ncol)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


def test_gather_columns_rebuilds_the_block_partition():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
    x = torch.ones(3, 7)
    assert gather_columns(x, 7) is x                            # no process group: identity


def test_ranks_are_spread_over_the_gpus_of_a_node(monkeypatch):
    """More GPUs than ranks: constant stride (the ranks do not share a host bridge); as many GPUs as ranks: identity;
    NOAHMP_GPU_MAP overrides."""
    from noahmp_b200.parallel import device_for_rank
    monkeypatch.delenv("NOAHMP_GPU_MAP", raising=False)
    assert [device_for_rank(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]
    assert [device_for_rank(r, 2, 8) for r in range(2)] == [0, 4]
    assert [device_for_rank(r, 8, 8) for r in range(8)] == list(range(8))
    assert [device_for_rank(r, 4, 4) for r in range(4)] == [0, 1, 2, 3]
    assert [device_for_rank(r, 2, 3) for r in range(2)] == [0, 1] and device_for_rank(0, 1, 8) == 0
    monkeypatch.setenv("NOAHMP_GPU_MAP", "4,5,6,7")
    assert [device_for_rank(r, 4, 8) for r in range(4)] == [4, 5, 6, 7]
