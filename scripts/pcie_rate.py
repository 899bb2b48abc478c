#!/usr/bin/env python
"""Pinned host<->device copy rates on this box (the floor of the NumPy-in / NumPy-out path)."""
import time

import torch

n = 614_400_000
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn in (("H2D", h2d), ("D2H", d2h), ("H2D+D2H concurrent", both)):
    t = timed(fn)
    print(f"{name}: {t * 1e3:.2f} ms for 614.4 MB each -> {n / t / 1e9:.1f} GB/s per direction")
