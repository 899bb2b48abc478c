"""PARRM.filter_data() on a PAGEABLE cfg2 recording (staged through pinned buffers by copy
threads): ms per pass for the copy-thread settings in the environment
(PYPARRM_B200_COPY_THREADS), next to the pinned path."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pyparrm_b200 import PARRM, pinned_empty, _engine
from pyparrm_b200.synthetic import make_recording, true_period

rec = make_recording(64, 1_200_000, 2000, 130, seed=0)
def rate(data, n=12):
    p = PARRM(data, 2000, 130, verbose=False)
    p._period = np.float64(true_period(2000, 130))
    p.create_filter(filter_half_width=2000, filter_direction="both")
    for _ in range(3):
        p.filter_data()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        p.filter_data()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / n * 1e3, 2)
print("copy threads", _engine._COPY_THREADS, "pageable ms", rate(rec), flush=True)
if os.environ.get("WITH_PINNED"):
    pin = pinned_empty(rec.shape); pin[...] = rec
    print("pinned ms", rate(pin))
