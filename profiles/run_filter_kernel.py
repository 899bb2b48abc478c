"""Short driver for ncu: a few device-resident passes of a named filter shape (default cfg2:
64 x 1.2M f64, 160 taps) through the kernel the library picks for a job of that size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyparrm_b200 import _engine  # noqa: E402

n_chans = int(os.environ.get("PROF_CHANS", "64"))
n_samples = int(os.environ.get("PROF_SAMPLES", "1200000"))
fs, fa, hw = (float(os.environ.get(k, d)) for k, d in
              (("PROF_FS", "2000"), ("PROF_FA", "130"), ("PROF_HW", "2000")))
engine = _engine.get_engine()
period = fs / fa * (1 + 3e-6)
taps = engine.build_taps(period, period / 50, int(hw), 0, os.environ.get("PROF_DIRECTION", "both"))
gen = torch.Generator(device="cuda").manual_seed(0)
d_x = torch.randn((n_chans, n_samples), dtype=torch.float64, device="cuda", generator=gen)
d_y = torch.empty_like(d_x)
kernel = os.environ.get("PROF_KERNEL")
tuning = {"variant": int(os.environ["PROF_VARIANT"])} if "PROF_VARIANT" in os.environ else None
for _ in range(int(os.environ.get("PROF_PASSES", "6"))):
    engine.filter_device(d_x, taps, d_out=d_y, kernel=None if kernel is None else int(kernel),
                         tuning=tuning)
torch.cuda.synchronize()
print("taps", taps.shape[0], "kernel", engine.last_filter_kernel, "checksum", float(d_y.sum()))
