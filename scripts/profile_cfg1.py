"""cProfile of the bundled-example workflow (cfg1) and of one cfg2 filter_data() pass."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pyparrm_b200 import PARRM, get_example_data_paths, pinned_empty
from pyparrm_b200.synthetic import make_recording, true_period

data = np.load(get_example_data_paths("example_data"))
def workflow():
    p = PARRM(data, 200, 150, verbose=False)
    p.find_period()
    p.create_filter(2000, 20, "both", 0.01)
    return p.filter_data()
for _ in range(3):
    t0 = time.perf_counter(); workflow(); print("cfg1 workflow", round((time.perf_counter() - t0) * 1e3, 2), "ms")
pr = cProfile.Profile(); pr.enable(); workflow(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

rec = pinned_empty((64, 1_200_000)); make_recording(64, 1_200_000, 2000, 130, seed=0, out=rec)
q = PARRM(rec, 2000, 130, verbose=False); q._period = np.float64(true_period(2000, 130))
q.create_filter(filter_half_width=2000)
for _ in range(3):
    q.filter_data()
pr = cProfile.Profile(); pr.enable(); q.filter_data(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
