"""Does the NVML clock sampler of bench.py disturb the end-to-end pass?  Times
PARRM.filter_data() on the cfg2 recording (pinned) with the sampler off / on at several periods."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from pyparrm_b200 import PARRM, pinned_empty
from pyparrm_b200.synthetic import make_recording, true_period

rec = pinned_empty((64, 1_200_000))
make_recording(64, 1_200_000, 2000, 130, seed=0, out=rec)
p = PARRM(rec, 2000, 130, verbose=False)
p._period = np.float64(true_period(2000, 130))
p.create_filter(filter_half_width=2000, filter_direction="both")
for _ in range(3):
    p.filter_data()

def run(n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        out = p.filter_data()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / n * 1e3, 2)

for rep in range(2):
    print("no sampler", run(), flush=True)
    for period in (0.002, 0.02, 0.1):
        s = bench.ClockSampler(0)
        s.period = period
        with s:
            print("sampler every", period, "s:", run(), "samples", len(s.samples), flush=True)
