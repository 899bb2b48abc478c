#!/usr/bin/env python
"""Per-phase cycle split of the strip kernel (debug build from scripts/build_timing.sh)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyparrm_b200 import _native  # noqa: E402

_native.LIB_PATH = os.environ.get("PARRM_TIMING_LIB", os.path.join(ROOT, "build", "libparrm_b200_timing.so"))
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
names = ["issue", "dpass(own)", "bar1", "gather", "wait+ctr", "bar2", "-", "-"]
pipe_names = ["g:wait_full", "g:gather", "g:arrive", "-", "s:wait_empty", "s:issue+slide", "s:wait_tma",
              "s:bar+arrive"]
for shape in sys.argv[1:] or ["512,4,2048,3,1"]:
    parts = shape.split(",")
    pipe = len(parts) == 6
    th, ru, tile, pre, ctas = parts[:5]
    os.environ.update(PARRM_FILTER_TILE=tile, PARRM_FILTER_THREADS=th, PARRM_FILTER_RU=ru,
                      PARRM_FILTER_PREFETCH=pre, PARRM_FILTER_CTAS=ctas,
                      PARRM_FILTER_PIPE="1" if pipe else "0",
                      PARRM_FILTER_SLIDE=parts[5] if pipe else "0")
    eng.filter_device(d_x, taps, d_out=d_y)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8)()
    _native.lib.parrm_debug_strip_timing(buf, 1)
    eng.filter_device(d_x, taps, d_out=d_y)
    torch.cuda.synchronize()
    _native.lib.parrm_debug_strip_timing(buf, 1)
    vals = np.array(list(buf), dtype=np.float64)
    n_ctas = 148 * int(ctas)
    steps = (C * ((T + 2 * int(tile) - 2) // int(tile))) / n_ctas
    print(shape, "steps/CTA ~%.0f" % steps,
          {n: round(v / steps) for n, v in zip(pipe_names if pipe else names, vals) if n != "-"},
          "total/step %.0f" % (vals.sum() / steps))
