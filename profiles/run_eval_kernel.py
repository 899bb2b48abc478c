"""Short driver for ncu: the candidate evaluator on the cfg2 run-3 shape (24.7k x 64, bw 20)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyparrm_b200 import _engine  # noqa: E402
from pyparrm_b200.synthetic import make_recording  # noqa: E402

n_chans = int(os.environ.get("PROF_CHANS", "64"))
n_cand = int(os.environ.get("PROF_CANDIDATES", "381"))
engine = _engine.get_engine()
data = make_recording(n_chans, 60_000, 2000, 130, seed=0)
idx = np.unique(np.random.default_rng(0).integers(0, 57_000, 25_000)) + 1500
(tile,) = engine.prepare_tiles(data, [idx], 3.0)
periods = 2000 / 130 * (1 + np.linspace(-3e-3, 3e-3, n_cand))
for _ in range(int(os.environ.get("PROF_PASSES", "4"))):
    err = engine.evaluate(tile, periods, 20, 1.0, n_chans)
print("best", periods[err.argmin()], err.min())

if os.environ.get("PROF_TABLE") == "1":  # kernel durations without ncu, for cross-checking
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        err = engine.evaluate(tile, periods, 20, 1.0, n_chans)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(f"  {e.key[:70]:70s} x{e.count} {e.device_time_total / e.count:.1f} us each")
