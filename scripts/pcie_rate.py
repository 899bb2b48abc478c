#!/usr/bin/env python
"""Pinned host<->device copy rates on this box: the floor of the NumPy-in / NumPy-out path.

Single process:  python scripts/pcie_rate.py
All GPUs at once (what N ranks of bench.py's e2e leg compete for: host memory and the PCIe
roots):  python -m torch.distributed.run --nproc-per-node N scripts/pcie_rate.py
Rank 0 prints one JSON line; with N ranks the times are the max over ranks and the rates the
aggregate over all ranks."""
import json
import os
import time

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 614_400_000  # one cfg2 recording
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    seconds = (time.perf_counter() - t0) / reps
    if dist is not None:
        t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds = float(t.item())
    return seconds


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


result = {"n_gpus": world, "bytes_per_direction_per_gpu": n}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("h2d_d2h_concurrent", both)):
    t = timed(fn)
    result[name] = {"ms": round(t * 1e3, 2), "aggregate_gbs_per_direction": round(world * n / t / 1e9, 1)}
# what the e2e leg could reach at best: one recording up and one down per GPU per pass
result["e2e_floor_channel_samples_per_s"] = world * (n / 8) / (result["h2d_d2h_concurrent"]["ms"] * 1e-3)
if int(os.environ.get("RANK", "0")) == 0:
    print(json.dumps(result), flush=True)
if dist is not None:
    dist.destroy_process_group()
