#!/usr/bin/env python
"""Launch time of parrm_filter_apply on cfg2 with an alternative build of the library."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyparrm_b200 import _native  # noqa: E402

_native.LIB_PATH = sys.argv[1]
_native.lib = _native._load()
from pyparrm_b200 import _engine  # noqa: E402

_engine.lib = _native.lib
from oracle import parrm_oracle as oracle  # noqa: E402

C, T = 64, 1_200_000
per = 2000 / 130 * (1 + 3e-6)
taps = oracle.tap_offsets(per, per / 50, 2000, 0, "both")
eng = _engine.get_engine()
d_x = torch.randn((C, T), dtype=torch.float64, device="cuda")
d_y = torch.empty_like(d_x)
for _ in range(3):
    eng.filter_device(d_x, taps, d_out=d_y)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    eng.filter_device(d_x, taps, d_out=d_y)
e1.record()
torch.cuda.synchronize()
print(os.path.basename(sys.argv[1]), "%.4f ms" % (e0.elapsed_time(e1) / 20))
