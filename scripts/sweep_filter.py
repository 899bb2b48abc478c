#!/usr/bin/env python
"""Time parrm_filter_apply on cfg2 (64 x 1.2M f64, 160 taps) for several strip-kernel shapes.

Tuning aid (run on a B200): shapes are forced through the PARRM_FILTER_* environment
variables that launch_strip() reads at every launch.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parrm_oracle as oracle  # noqa: E402  (parity check of every shape)
from pyparrm_b200 import _engine, _native  # noqa: E402
from pyparrm_b200.synthetic import make_recording  # noqa: E402

C, T, FS, FA = 64, 1_200_000, 2000, 130
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if cfg == "cfg4":
    C, T, FS, FA = 48, 2_000_000, 30000, 130
x = make_recording(C, T, FS, FA, seed=0)
per = FS / FA * (1 + 3e-6)
if cfg == "cfg4":
    taps = oracle.tap_offsets(per, per / 50, 2311, 0, "past")
else:
    taps = oracle.tap_offsets(per, per / 50, 2000, 0, "both")
eng = _engine.get_engine()
d_x = torch.from_numpy(x).cuda()
d_y = torch.empty_like(d_x)
want = oracle.apply_filter_direct(x[:2], taps)
scale = np.abs(x).max()


def run(label, strategy, env):
    for k in list(os.environ):
        if k.startswith("PARRM_FILTER_"):
            del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    d_y.zero_()
    for _ in range(3):
        eng.filter_device(d_x, taps, d_out=d_y, strategy=strategy)
    torch.cuda.synchronize()
    err = float(np.abs(d_y[:2].cpu().numpy() - want).max() / scale)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        eng.filter_device(d_x, taps, d_out=d_y, strategy=strategy)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gbs = 16.0 * C * T / (ms * 1e-3) / 1e9
    print(json.dumps({"shape": label, "ms": round(ms, 4), "GB/s": round(gbs, 1),
                      "frac_of_6549": round(gbs / 6549.1, 4), "rel_err": err}), flush=True)


run("gather", _native.PLAN_GATHER, {})
run("auto-default", _native.PLAN_AUTO, {})
for pipe, threads, slide, ru, tile, pre, ctas, stages in [
    (1, 512, 512, 4, 2048, 1, 1, 2),
    (1, 256, 256, 6, 1536, 1, 1, 3), (1, 512, 512, 3, 1536, 1, 1, 3),
    (1, 320, 320, 4, 1280, 1, 1, 3), (1, 640, 256, 2, 1280, 1, 1, 3), (1, 256, 256, 5, 1280, 1, 1, 3),
    (1, 512, 512, 2, 1024, 1, 1, 3), (1, 256, 256, 4, 1024, 2, 1, 3), (1, 512, 512, 2, 1024, 1, 1, 4),
]:
    run(f"pipe{pipe} t{threads}+{slide} ru{ru} tile{tile} pre{pre} ctas{ctas} stages{stages}",
        _native.PLAN_AUTO,
        {"PARRM_FILTER_TILE": tile, "PARRM_FILTER_THREADS": threads, "PARRM_FILTER_RU": ru,
         "PARRM_FILTER_PREFETCH": pre, "PARRM_FILTER_CTAS": ctas, "PARRM_FILTER_PIPE": pipe,
         "PARRM_FILTER_SLIDE": slide, "PARRM_FILTER_STAGES": stages})
