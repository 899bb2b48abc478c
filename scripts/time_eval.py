#!/usr/bin/env python
"""Device time of one batched evaluation (bench.py's find_period shape) with SM clock samples."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyparrm_b200 import _native  # noqa: E402

if os.environ.get("PARRM_TIMING_LIB"):  # alternative build of the library (timing experiments)
    _native.LIB_PATH = os.environ["PARRM_TIMING_LIB"]
    _native.lib = _native._load()
from pyparrm_b200 import _engine  # noqa: E402

_engine.lib = _native.lib
from pyparrm_b200.synthetic import make_recording  # noqa: E402

n_chans = int(os.environ.get("EVAL_CHANS", "64"))
n_cand = int(os.environ.get("EVAL_CANDIDATES", "3048"))
engine = _engine.get_engine()
n_samples = int(os.environ.get("EVAL_SAMPLES", "1200000"))
data = make_recording(n_chans, n_samples, 2000, 130, seed=0)
rng = np.random.default_rng(0)
lo, hi = int(0.025 * n_samples), int(0.975 * n_samples)
if os.environ.get("EVAL_IDX", "large") == "small":
    lo, hi = 1_000, 58_000
idx = np.unique(rng.integers(0, hi - lo, int(os.environ.get("EVAL_DRAWS", "25000")))) + lo
(tile,) = engine.prepare_tiles(data, [idx], 3.0)
periods = 2000 / 130 * (1 + np.linspace(-3e-3, 3e-3, n_cand))
d_per = torch.from_numpy(periods).cuda()
for _ in range(2):
    engine.evaluate_device(tile, d_per, 20, 1.0, n_chans)
torch.cuda.synchronize()
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active",
                        "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(0.2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    engine.evaluate_device(tile, d_per, 20, 1.0, n_chans)
e1.record()
torch.cuda.synchronize()
mon.terminate()
lines = mon.stdout.read().strip().splitlines()
ms = e0.elapsed_time(e1) / reps
print(f"{n_cand} candidates x {len(idx)} samples x {n_chans} ch: {ms:.3f} ms per call, "
      f"{n_cand / ms * 1e3:.0f} cand/s, {ms / n_cand * 1e3:.2f} us per candidate")
if len(lines) >= 4:
    print("clock samples (MHz, W, reasons):", lines[len(lines) // 4], "|", lines[len(lines) // 2], "|", lines[-2])

if hasattr(_native.lib, "parrm_debug_tensor_timing"):  # -DPARRM_TENSOR_TIMING build: phase split
    import ctypes

    buf = (ctypes.c_ulonglong * 128)()
    _native.lib.parrm_debug_tensor_timing(buf, 1)
    engine.evaluate_device(tile, d_per, 20, 1.0, n_chans)
    torch.cuda.synchronize()
    _native.lib.parrm_debug_tensor_timing(buf, 0)
    names = ["sincos", "sync", "generate", "stage_y", "cp_wait", "sync", "multiply", "sync"]
    tiles = -(-len(idx) // 128)
    print("cycles per 128-sample tile, lane 0 of each warp (barrier waits show up in the next phase):")
    for w in range(16):
        vals = [buf[w * 8 + i] for i in range(8)]
        print(f"  warp {w:2d}", {n + str(i): round(v / tiles) for i, (n, v) in enumerate(zip(names, vals))},
              "total", round(sum(vals) / tiles))

if hasattr(_native.lib, "parrm_debug_solve_timing"):  # -DPARRM_SOLVE_TIMING build: phase split
    import ctypes

    buf = (ctypes.c_ulonglong * 8)()
    _native.lib.parrm_debug_solve_timing(buf, 1)
    engine.evaluate_device(tile, d_per, 20, 1.0, n_chans)
    torch.cuda.synchronize()
    _native.lib.parrm_debug_solve_timing(buf, 0)
    names = ["gram", "lu", "load_b", "forward", "back", "beta+quad"]
    print("solve kernel, cycles of thread 0 of CTA 0:", {n: int(buf[i]) for i, n in enumerate(names)},
          "total", sum(int(buf[i]) for i in range(6)))

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        engine.evaluate_device(tile, d_per, 20, 1.0, n_chans)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages()
        if e.device_time_total > 0]
for key, count, us in sorted(rows, key=lambda r: -r[2])[:6]:
    print(f"  {key[:60]:60s} x{count}  {us:.1f} us each")
