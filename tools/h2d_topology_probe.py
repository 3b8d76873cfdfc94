#!/usr/bin/env python3
"""Which GPUs share a host link?  Pinned host-to-device bandwidth with W ranks copying at once, for several rank -> GPU maps
(MEASUREMENT TOOL; torchrun --nproc-per-node W tools/h2d_topology_probe.py).  On a box whose GPUs hang off several PCIe root
ports a placement that spreads the ranks over the ports gets more host bandwidth than the first W GPUs do."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

MAPS = {2: [[0, 1], [0, 2], [0, 4], [0, 7]], 4: [[0, 1, 2, 3], [0, 2, 4, 6], [0, 1, 4, 5], [4, 5, 6, 7]], 8: [list(range(8))],
        1: [[0], [4]]}


def main():
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        dist.init_process_group("gloo")
    nbytes, reps = 1 << 30, 6
    h = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    h.fill_(1.0)
    out = []
    for m in MAPS.get(world, [list(range(world))]):
        if max(m) >= torch.cuda.device_count():
            continue
        dev = torch.device("cuda", m[rank])
        d = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(dev)
        mine = time.perf_counter() - t0
        el = torch.tensor([mine], dtype=torch.float64)
        each = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        if world > 1:
            dist.all_gather(each, el.clone())
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        else:
            each = [el.clone()]
        gbs = nbytes * reps / float(el.item()) / 1e9
        out.append({"map": m, "gbs_per_gpu": gbs, "gbs_all": gbs * world,
                    "gbs_each_rank_own_time": [round(nbytes * reps / float(t.item()) / 1e9, 1) for t in each]})
        del d
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"ranks": world, "results": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
