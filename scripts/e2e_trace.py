"""Per-pass wall times of PARRM.filter_data() on the cfg2 recording (pinned input), in the
order bench.py runs things, to find out where its occasional slow e2e window comes from."""
import gc, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from pyparrm_b200 import PARRM, _engine, pinned_empty
from pyparrm_b200.synthetic import make_recording, true_period

engine = _engine.get_engine()
rec = pinned_empty((64, 1_200_000))
make_recording(64, 1_200_000, 2000, 130, seed=0, out=rec)
p = PARRM(rec, 2000, 130, verbose=False)
p._period = np.float64(true_period(2000, 130))
p.create_filter(filter_half_width=2000, filter_direction="both")
taps = (np.flatnonzero(p.filter < 0) - 2000).astype(np.int32)
d_x = torch.from_numpy(rec).cuda(); d_y = torch.empty_like(d_x)
mode = sys.argv[1] if len(sys.argv) > 1 else "bench"
if mode == "bench":
    for _ in range(55):
        engine.filter_device(d_x, taps, d_out=d_y)
    torch.cuda.synchronize()
    bench.bench_standardise(engine, d_x, 6549.1)
for _ in range(3):
    p.filter_data()
times = []
gcs = []
def cb(phase, info):
    if phase == "start":
        gcs.append((len(times), info["generation"], time.perf_counter()))
    else:
        i, g, t0 = gcs[-1]
        gcs[-1] = (i, g, round((time.perf_counter() - t0) * 1e3, 2))
gc.callbacks.append(cb)
for i in range(80):
    t0 = time.perf_counter()
    out = p.filter_data()
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
print(mode, "median %.2f ms" % np.median(times), "slow passes (>18 ms):",
      [(i, round(t, 1)) for i, t in enumerate(times) if t > 18])
print("gc events (pass, generation, ms):", gcs[:20])
print("pinned live MB", getattr(engine, "_pinned_live", 0) / 2**20)
