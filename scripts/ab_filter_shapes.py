import json, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import parrm_oracle as oracle
from pyparrm_b200 import _native as K
from pyparrm_b200._engine import get_engine
eng = get_engine()
def taps_of(fs, fa, hw, d):
    p = fs / fa * (1 + 3e-6)
    return oracle.tap_offsets(p, p / 50, hw, 0, d)
cases = {"cfg3": taps_of(1000, 145, 2469, "both"), "cfg4": taps_of(30000, 130, 2311, "past"), "cfg2": taps_of(2000, 130, 2000, "both")}
shapes = {"cfg3": [(64, 1_200_000), (256, 3_600_000)], "cfg4": [(64, 1_200_000), (384, 3_000_000)], "cfg2": [(64, 1_200_000)]}
cases["cfg1"] = oracle.tap_offsets(1.3311148014466094, 0.01, 2000, 20, "both")
shapes["cfg1"] = [(64, 1_200_000)]
VARIANTS = [{}, {"variant": 4}]  # 4 = strips of equal length (the round-2 starting point)
tun = {name: VARIANTS for name in ("cfg1", "cfg2", "cfg3", "cfg4")}
for name, taps in cases.items():
    for (c, n) in shapes[name]:
        d_x = torch.randn((c, n), dtype=torch.float64, device="cuda")
        d_y = torch.empty_like(d_x)
        res = {json.dumps(t): [] for t in tun[name]}
        for rep in range(3):
            for t in tun[name]:
                try:
                    for _ in range(2):
                        eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
                    torch.cuda.synchronize()
                except RuntimeError:
                    res[json.dumps(t)].append(None)
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                res[json.dumps(t)].append(round(16 * c * n / ms / 1e6 / 6549.1, 4))
        print(name, (c, n), res, flush=True)
        del d_x, d_y
