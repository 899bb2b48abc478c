"""Short driver for ncu: a few device-resident passes of the cfg2 filter (64 x 1.2M f64, 160 taps)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyparrm_b200 import _engine  # noqa: E402

n_chans = int(os.environ.get("PROF_CHANS", "64"))
n_samples = int(os.environ.get("PROF_SAMPLES", "1200000"))
engine = _engine.get_engine()
period = 2000 / 130 * (1 + 3e-6)
taps = engine.build_taps(period, period / 50, 2000, 0, os.environ.get("PROF_DIRECTION", "both"))
gen = torch.Generator(device="cuda").manual_seed(0)
d_x = torch.randn((n_chans, n_samples), dtype=torch.float64, device="cuda", generator=gen)
d_y = torch.empty_like(d_x)
for _ in range(int(os.environ.get("PROF_PASSES", "6"))):
    engine.filter_device(d_x, taps, d_out=d_y)
torch.cuda.synchronize()
print("taps", taps.shape[0], "checksum", float(d_y.sum()))
