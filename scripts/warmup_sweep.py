"""Device-resident cfg2 filter pass after an idle second, as a function of the warm-up
duration before the timed region (20 passes): too short and the GPU is not at its boost
clocks yet, too long and it sits at its power cap."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import parrm_oracle as oracle
from pyparrm_b200 import _engine
eng = _engine.get_engine()
period = 2000 / 130 * (1 + 3e-6)
taps = oracle.tap_offsets(period, period / 50, 2000, 0, "both")
d_x = torch.randn((64, 1_200_000), dtype=torch.float64, device="cuda"); d_y = torch.empty_like(d_x)
for _ in range(5):
    eng.filter_device(d_x, taps, d_out=d_y)
torch.cuda.synchronize()
for rep in range(2):
    for warm_ms in (0, 2, 5, 10, 20, 50, 100, 200, 500):
        time.sleep(1.0)
        t0 = time.perf_counter()
        n_warm = 0
        while (time.perf_counter() - t0) * 1e3 < warm_ms:
            for _ in range(4):
                eng.filter_device(d_x, taps, d_out=d_y)
            torch.cuda.synchronize(); n_warm += 4
        res = []
        for steps in (20, 50):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                eng.filter_device(d_x, taps, d_out=d_y)
            e1.record(); torch.cuda.synchronize()
            res.append(round(e0.elapsed_time(e1) / steps, 4))
        print(f"warm-up {warm_ms:4d} ms ({n_warm:4d} passes): 20-step region {res[0]} ms/pass, next 50-step region {res[1]} ms/pass", flush=True)
