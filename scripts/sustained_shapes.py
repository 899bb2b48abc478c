"""cfg2 filter pass at the GPU's power cap (after 0.5 s of back-to-back passes) and in the
burst state (10 ms warm-up after a 1 s pause), for a few launch shapes of the specialised
kernel: does a shape that draws less power win when the power cap, not the shared-memory
pipe, sets the clock?"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from oracle import parrm_oracle as oracle
from pyparrm_b200 import _native as K
from pyparrm_b200._engine import get_engine
eng = get_engine()
p = 2000 / 130 * (1 + 3e-6)
taps = oracle.tap_offsets(p, p / 50, 2000, 0, "both")
d_x = torch.randn((64, 1_200_000), dtype=torch.float64, device="cuda"); d_y = torch.empty_like(d_x)
def run(n, t):
    for _ in range(n):
        eng.filter_device(d_x, taps, d_out=d_y, kernel=K.KERNEL_SPECIALISED, tuning=t)
def timed(n, t):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(n, t); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = [{}, {"ctas_per_sm": 1}, {"steps_per_chunk": 5}, {"steps_per_chunk": 20}, {"ctas_per_sm": 1, "steps_per_chunk": 20},
          {"prefetch_chunks": 1}, {"prefetch_chunks": 4}]
for t in shapes:
    try:
        run(3, t); torch.cuda.synchronize()
    except RuntimeError as e:
        print(json.dumps(t), "not runnable"); continue
    res = {}
    for rep in range(2):
        time.sleep(1.0); run(40, t); torch.cuda.synchronize()
        res.setdefault("burst_ms", []).append(round(timed(20, t), 4))
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.5:
            run(40, t); torch.cuda.synchronize()
        res.setdefault("sustained_ms", []).append(round(timed(50, t), 4))
    print(json.dumps(t), res, flush=True)
