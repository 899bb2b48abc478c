import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import parrm_oracle as oracle
from pyparrm_b200 import _native as K
from pyparrm_b200._engine import get_engine
eng = get_engine()
def taps_of(fs, fa, hw, d):
    p = fs / fa * (1 + 3e-6)
    return oracle.tap_offsets(p, p / 50, hw, 0, d)
cases = {"cfg2": taps_of(2000, 130, 2000, "both"), "cfg3": taps_of(1000, 145, 2469, "both"), "cfg4": taps_of(30000, 130, 2311, "past")}
TRY = {"cfg2": (5, 10, 20), "cfg3": (5, 25), "cfg4": (10, 20)}
shapes = {"cfg2": (64, 1_200_000), "cfg3": (256, 3_600_000), "cfg4": (384, 3_000_000)}
for name, taps in cases.items():
    c, n = shapes[name]
    d_x = torch.randn((c, n), dtype=torch.float32, device="cuda"); d_y = torch.empty_like(d_x)
    grid = [{}] + [{"ctas_per_sm": c, "steps_per_chunk": u} for c in (2, 3, 4) for u in TRY[name]]
    for t in grid:
        try:
            for _ in range(3):
                eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
            torch.cuda.synchronize()
        except RuntimeError as e:
            print(name, t, "ERR", str(e)[:80]); continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(name, t, eng.last_filter_kernel, "ms %.4f  G samples/s %.1f  frac of 8B roofline %.3f" % (ms, c * n / ms / 1e6, 8 * c * n / ms / 1e6 / 6549.1), flush=True)
    x = d_x[:2, :50000].double().cpu().numpy()
    got = d_y[:2, 2500:47000].cpu().numpy()
    want = oracle.apply_filter_direct(d_x[:2, :50000].double().cpu().numpy(), taps)[:, 2500:47000]
    print("  err", np.abs(got - want).max() / np.abs(x).max())
