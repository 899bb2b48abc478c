#!/usr/bin/env python
"""BASELINE cfg5: the evaluator sweep P in {1e4, 1e5, 1e6} candidate periods x N in {1e4, 1e5,
1e6} search samples (bandwidth 20, lambda 1, one channel, contiguous indices), candidates in
contiguous blocks per GPU, winner by one (error, index) pair per rank.
python scripts/cfg5_grid.py            (one GPU)
python -m torch.distributed.run --nproc-per-node N scripts/cfg5_grid.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parrm_oracle as oracle  # noqa: E402  (standardise of the synthetic trace only)
from pyparrm_b200 import _engine, _sharding, enable_sharding  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    enable_sharding(gather="none")
engine = _engine.get_engine()
fs, fa = 2000, 130
p_true = fs / fa * (1 + 3e-6)
rng = np.random.default_rng(7)
n_max = 1_000_000
t = np.arange(n_max + 1, dtype=np.float64)
trace = (np.sin(2 * np.pi * t / p_true) + 0.5 * rng.standard_normal(n_max + 1))[None, :]
z = oracle.standardise(trace, 3.0)
rows = []
for n_fit in (10_000, 100_000, 1_000_000):
    tile = engine.tile_from_standardised(z, np.arange(n_fit))
    for n_cand in (10_000, 100_000, 1_000_000):
        sweep = (fs / fa) * (1 + np.linspace(-1e-2, 1e-2, n_cand))
        evaluate = lambda blk, tile=tile: engine.evaluate_device(tile, blk, 20, 1.0, 1)  # noqa: E731
        evaluate(sweep[:1024])
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        if dist is not None:
            best, err = _sharding.minloc_sharded(evaluate, sweep)
        else:
            err, best = engine.argmin(evaluate(sweep))
        torch.cuda.synchronize()
        seconds = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([seconds], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            seconds = float(tt.item())
        rows.append({"n_gpus": world, "candidates": n_cand, "samples": n_fit, "seconds": round(seconds, 4),
                     "candidates_per_s": n_cand / seconds,
                     "sample_candidates_per_s": n_cand * n_fit / seconds,
                     "winner_rel_err_vs_injected": abs(float(sweep[best]) - p_true) / p_true})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
if dist is not None:
    dist.destroy_process_group()
