"""Per-CTA timeline of the specialised filter kernel (profiling aid: options.variant bit 1 +
options.timeline): SM id, start / end in ns, pieces per CTA.  python scripts/filter_timeline.py [cfg]"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import parrm_oracle as oracle
from pyparrm_b200 import _native as K
from pyparrm_b200._engine import get_engine

eng = get_engine()
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
fs, fa, hw, d, c, n = {"cfg2": (2000, 130, 2000, "both", 64, 1_200_000),
                       "cfg3": (1000, 145, 2469, "both", 256, 3_600_000),
                       "cfg4": (30000, 130, 2311, "past", 384, 3_000_000)}[name]
p = fs / fa * (1 + 3e-6)
taps = oracle.tap_offsets(p, p / 50, hw, 0, d)
d_x = torch.randn((c, n), dtype=torch.float64, device="cuda")
d_y = torch.empty_like(d_x)
tl = torch.zeros(4 * 1024, dtype=torch.int64, device="cuda")
tun = {"variant": 2 | (int(sys.argv[2]) if len(sys.argv) > 2 else 0), "timeline": tl.data_ptr()}
for _ in range(3):
    eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=tun)
torch.cuda.synchronize()
t = tl.cpu().numpy().reshape(-1, 4)
t = t[t[:, 2] > 0]
t0 = t[:, 1].min()
dur = (t[:, 2] - t[:, 1]) / 1e3
print(name, "CTAs", len(t), "span us", (t[:, 2].max() - t0) / 1e3)
print("start us: min %.1f max %.1f" % ((t[:, 1].min() - t0) / 1e3, (t[:, 1].max() - t0) / 1e3))
print("end   us: min %.1f p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile((t[:, 2] - t0) / 1e3, [0, 50, 90, 100])))
print("dur   us: min %.1f p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile(dur, [0, 50, 90, 100])))
for k in sorted(set(t[:, 3])):
    m = t[:, 3] == k
    print("pieces", k, "n", m.sum(), "dur mean %.1f" % dur[m].mean())
sm = t[:, 0]
per_sm = {}
for s, e in zip(sm, (t[:, 2] - t0) / 1e3):
    per_sm[s] = max(per_sm.get(s, 0), e)
ends = np.array([per_sm[s] for s in sorted(per_sm)])
print("SMs", len(per_sm), "CTAs per SM", np.bincount(np.bincount(sm.astype(int))))
print("per-SM end us by SM id (every 8th):", np.round(ends[::8], 1).tolist())
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/timeline_{name}.npy", t)
