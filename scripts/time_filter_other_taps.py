#!/usr/bin/env python
"""Device-resident filter pass (64 x 1.2 M float64) for tap structures away from the BASELINE
configs -- the bundled example with create_filter()'s defaults, a long-window comb on a stride
of 537, runs of consecutive taps -- with the kernel the library picks and with the pre-built
tap-by-tap gather forced (what these jobs ran on before the planner bounded its box lengths
by the specialised kernel's registers).  Timing only; parity is tests/test_gpu_filter.py."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parrm_oracle as oracle  # noqa: E402  (tap sets only)
from pyparrm_b200 import _engine, _native  # noqa: E402

CASES = {
    "example default (period 1.33, hw 2408)": (1.3311148014466094, None, 2408, 0, "both"),
    "ecog-like (period 8.66, hw 5000)": (8.6613, None, 5000, 0, "both"),
    "wide runs (period 15.38, phw 1.0)": (2000 / 130, 1.0, 2000, 0, "both"),
    "wide runs past (period 230.8, phw 20, hw 5000)": (30000 / 130, 20.0, 5000, 0, "past"),
    "cfg1 taps (period 1.33, phw .01, omit 20)": (1.3311148014466094, 0.01, 2000, 20, "both"),
}
eng = _engine.get_engine()
d_x = torch.randn((64, 1_200_000), dtype=torch.float64, device="cuda")
d_y = torch.empty_like(d_x)
for name, (period, phw, hw, omit, direction) in CASES.items():
    taps = oracle.tap_offsets(period, period / 50 if phw is None else phw, hw, omit, direction)
    row = {"taps": name, "n_taps": int(len(taps))}
    for label, kernel in (("auto", None), ("gather", _native.KERNEL_GATHER)):
        reps = 10 if label == "auto" else 2
        for _ in range(2):
            eng.filter_device(d_x, taps, d_out=d_y, kernel=kernel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.filter_device(d_x, taps, d_out=d_y, kernel=kernel)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row[label] = {"kernel": eng.last_filter_kernel, "ms": round(ms, 3),
                      "frac_of_hbm_6549": round(16.0 * d_x.numel() / (ms * 1e-3) / 1e9 / 6549.1, 3)}
    print(json.dumps(row), flush=True)
