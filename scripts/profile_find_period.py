"""Where the wall time of PARRM.find_period() goes on the cfg2 recording (cProfile, host side;
device waits show up in the call that synchronises)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pyparrm_b200 import PARRM
from pyparrm_b200.synthetic import make_recording

data = make_recording(64, 1_200_000, 2000, 130, seed=0)
p = PARRM(data, 2000, 130, verbose=False)
p.find_period(random_seed=0)
for _ in range(2):
    t0 = time.perf_counter(); p.find_period(random_seed=0); print("find_period", round((time.perf_counter() - t0) * 1e3, 1), "ms")
pr = cProfile.Profile(); pr.enable(); p.find_period(random_seed=0); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
