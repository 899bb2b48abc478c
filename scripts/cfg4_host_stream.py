#!/usr/bin/env python
"""BASELINE cfg4 through the host API at the per-GPU share of an 8-GPU job: 48 channels x
9.0 M samples (30 kHz x 300 s), one-sided default-width filter, streamed through
PARRM.filter_data() in time chunks (a row is 72 MB, larger than the 32 MB ring chunks).
Reports float64-in/float64-out, and int16-in (as such probes deliver it) / float32-out.
python scripts/cfg4_host_stream.py [n_chans]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parrm_oracle as oracle  # noqa: E402  (spot parity only)
from pyparrm_b200 import PARRM, _engine  # noqa: E402

n_chans = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n, fs, fa = 9_000_000, 30000, 130
rng = np.random.default_rng(0)
period = fs / fa * (1 + 3e-6)
t = np.arange(n)
wave = sum(np.sin(2 * np.pi * k * t / period + k) / k for k in range(1, 6))
x16 = np.empty((n_chans, n), dtype=np.int16)
for c in range(n_chans):
    x16[c] = np.clip(np.round(200 * rng.standard_normal(n) + 600 * wave), -32768, 32767)
rows = {}
for name, data, kw in (("float64 in, float64 out", x16.astype(np.float64), {}),
                       ("int16 in, float32 out", x16, {"out_dtype": np.float32})):
    parrm = PARRM(data, fs, fa, verbose=False)
    parrm._period = np.float64(period)
    parrm.create_filter(filter_direction="past")
    taps = (np.flatnonzero(parrm.filter < 0) - parrm._filter_half_width).astype(np.int32)
    parrm.filter_data(**kw)
    t0 = time.perf_counter()
    out = parrm.filter_data(**kw)
    seconds = time.perf_counter() - t0
    lo = 4_000_000
    want = oracle.apply_filter_direct(x16[1:2, lo - 3000: lo + 8000].astype(np.float64), taps)[0, 3000:8000]
    err = float(np.abs(out[1, lo: lo + 5000] - want).max() / np.abs(x16).max())
    rows[name] = {"seconds": round(seconds, 4), "channel_samples_per_s": n_chans * n / seconds,
                  "gb_over_pcie": (data.nbytes + out.nbytes) / 1e9, "spot_rel_err": err,
                  "kernel": _engine.get_engine().last_filter_kernel, "taps": int(len(taps))}
print(json.dumps({"workload": f"cfg4 share: {n_chans} ch x {n} samples through PARRM.filter_data() "
                              "(time-chunked host pipeline)", "results": rows}))
