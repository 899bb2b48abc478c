#!/usr/bin/env python
"""Device-resident filter pass at the channel counts and tap sets of BASELINE.json's cfg2-cfg4.

The recordings are random data created on the device (timing only; parity at these tap sets is
tests/test_gpu_filter.py).  cfg3 and cfg4 are shortened in time -- the strip kernel's rate does
not depend on the recording length once every SM has strips -- so that the run stays small.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parrm_oracle as oracle  # noqa: E402  (tap sets only)
from pyparrm_b200 import _engine  # noqa: E402

SHAPES = [  # name, channels, samples, fs, artefact Hz, half width, direction
    ("cfg2 64ch 2kHz 130Hz both", 64, 1_200_000, 2000, 130, 2000, "both"),
    ("cfg3 256ch 1kHz 145Hz both (1.2M of 3.6M samples)", 256, 1_200_000, 1000, 145, 2469, "both"),
    ("cfg4 384ch 30kHz 130Hz past (1.5M of 9M samples)", 384, 1_500_000, 30000, 130, 2311, "past"),
]
eng = _engine.get_engine()
for name, n_chans, n_samples, fs, fa, hw, direction in SHAPES:
    period = fs / fa * (1 + 3e-6)
    taps = oracle.tap_offsets(period, period / 50, hw, 0, direction)
    d_x = torch.randn((n_chans, n_samples), dtype=torch.float64, device="cuda")
    d_y = torch.empty_like(d_x)
    for _ in range(3):
        eng.filter_device(d_x, taps, d_out=d_y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        eng.filter_device(d_x, taps, d_out=d_y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = 16.0 * n_chans * n_samples / (ms * 1e-3) / 1e9
    print(json.dumps({"shape": name, "taps": int(len(taps)), "ms": round(ms, 3),
                      "G channel-samples/s": round(n_chans * n_samples / ms / 1e6, 1),
                      "GB/s": round(gbs, 1), "frac_of_6549": round(gbs / 6549.1, 3)}), flush=True)
    del d_x, d_y
