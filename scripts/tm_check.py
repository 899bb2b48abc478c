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
cases = {"cfg2": taps_of(2000, 130, 2000, "both"), "cfg3": taps_of(1000, 145, 2469, "both"), "cfg4": taps_of(30000, 130, 2311, "past"),
         "cfg1": oracle.tap_offsets(1.3311148014466094, 0.01, 2000, 20, "both")}
rng = np.random.default_rng(1)
for v in ((16, 80) if '--parity' in sys.argv else ()):
    for name, taps in cases.items():
        for shape in [(3, 50_000), (2, 1999), (5, 20_011)]:
            x = rng.standard_normal(shape) * 3 + 10
            want = oracle.apply_filter_direct(x, taps)
            try:
                got = eng.filter_device(torch.from_numpy(x).cuda(), taps, kernel=K.KERNEL_SPECIALISED, tuning={"variant": v}).cpu().numpy()
            except RuntimeError as e:
                print(v, name, shape, "ERR", str(e)[:100]); continue
            print(v, name, shape, eng.last_filter_kernel, "err %.2e" % (np.abs(got - want).max() / np.abs(x).max()), flush=True)
    # non-finite
    taps = cases["cfg2"]
    x = rng.standard_normal((2, 40_000)); x[0, 12_345] = np.nan; x[1, 30_000] = np.inf; x[1, 5] = 1e12
    want = oracle.apply_filter_direct(x, taps); want[~np.isfinite(want)] = 0.0
    got = eng.filter_device(torch.from_numpy(x).cuda(), taps, kernel=K.KERNEL_SPECIALISED, tuning={"variant": v}).cpu().numpy()
    print(v, "non-finite: finite", np.isfinite(got).all(), "err %.2e" % (np.abs(got - want).max() / 1e12), "zeros", int((got == 0).sum()), int((want == 0).sum()), flush=True)
shapes = {"cfg2": (64, 1_200_000), "cfg3": (256, 3_600_000), "cfg4": (384, 3_000_000), "cfg1": (64, 1_200_000)}
TUN = [{}, {"variant": 80}]
for u, pf in ((5, 4), (5, 6), (5, 7), (4, 8)):
    TUN += [{"steps_per_chunk": u, "prefetch_chunks": pf}, {"variant": 80, "steps_per_chunk": u, "prefetch_chunks": pf}]
for name in sys.argv[1:] or list(cases):
    taps = cases[name]
    c, n = shapes[name]
    d_x = torch.randn((c, n), dtype=torch.float64, device="cuda"); d_y = torch.empty_like(d_x)
    res = {}
    for rep in range(2):
        for t in TUN:
            key = json.dumps(t)
            try:
                for _ in range(2):
                    eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
                torch.cuda.synchronize()
            except RuntimeError as e:
                res.setdefault(key, []).append(str(e)[:40]); continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
            e1.record(); torch.cuda.synchronize()
            res.setdefault(key, []).append(round(16 * c * n / (e0.elapsed_time(e1) / 10) / 1e6 / 6549.1, 4))
    print(name, (c, n))
    for k, v in res.items():
        print("   ", k, v, flush=True)
    del d_x, d_y
