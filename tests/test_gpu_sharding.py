"""The product's sharded API on real GPUs over NCCL (needs >= 2 GPUs; skipped on a 1-GPU box).

``enable_sharding()`` + ``PARRM.find_period()`` / ``PARRM.filter_data()`` on two ranks:
the period must be bit-identical to the unsharded one (and to the reference's golden value),
the filter shards must equal the oracle's direct sum, in every gather mode, for channel
shards, uneven channel blocks and time shards.  Run with ``gpurun --gpus 2``.
"""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_path):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import parrm_oracle as oracle
        from pyparrm_b200 import PARRM, _engine, _sharding, disable_sharding, enable_sharding
        from pyparrm_b200.synthetic import make_recording

        engine = _engine.get_engine()
        assert engine.device.index == rank
        checks = {}
        # ---- period: unsharded vs sharded, 3 channels (uneven blocks) ----------------------
        data = make_recording(3, 30_000, 200, 13, seed=5)
        p0 = PARRM(data, 200, 13, verbose=False)
        p0.find_period(random_seed=0)
        launches0 = engine.launches
        enable_sharding()
        p1 = PARRM(data, 200, 13, verbose=False)
        p1.find_period(random_seed=0)
        checks["period_bit_equal"] = float(p0.period == p1.period)
        checks["period"] = float(p1.period)
        checks["sharded_launches"] = float(engine.launches - launches0)
        # ---- filter: every gather mode, channel shards (uneven) and time shards ------------
        taps = oracle.tap_offsets(p1.period, p1.period / 50, 300, 0, "both")
        want = oracle.apply_filter_direct(data, taps)
        scale = np.abs(data).max()
        worst = 0.0
        for mode in ("rank0", "all", "none"):
            for rows in (slice(0, 3), slice(0, 1)):  # 3 channels -> blocks; 1 channel -> time shards
                enable_sharding(gather=mode)
                p = PARRM(data[rows], 200, 13, verbose=False)
                p._period = p1.period
                p.create_filter(300, 0, "both")
                got = p.filter_data()
                c0, c1, t0, t1 = p.filter_shard
                full = mode == "all" or (mode == "rank0" and rank == 0)
                ref = want[rows] if full else want[rows][c0:c1, t0:t1]
                assert got.shape == ref.shape, (mode, rows, got.shape, ref.shape)
                worst = max(worst, float(np.abs(got - ref).max() / scale))
        checks["filter_worst_rel"] = worst
        # ---- sweep winner with one (error, index) pair per rank ---------------------------
        idx = np.arange(100, 5100)
        z = oracle.standardise(data, 3.0)
        tile = engine.tile_from_standardised(z, idx)
        sweep = p1.period * (1 + np.linspace(-1e-3, 1e-3, 257))
        errs = engine.evaluate(tile, sweep, 5, 1.0, 3)
        best = _sharding.minloc_sharded(
            lambda blk: engine.evaluate_device(tile, blk, 5, 1.0, 3), sweep)
        checks["minloc_ok"] = float(best[0] == int(np.argmin(errs)) and
                                    abs(best[1] - errs.min()) <= 1e-12 * abs(errs.min()))
        disable_sharding()
        gathered = [None] * world
        dist.all_gather_object(gathered, checks)
        if rank == 0:
            import json

            with open(result_path, "w") as fh:
                json.dump(gathered, fh)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_ranks_nccl(tmp_path):
    import json

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world, port = 2, _free_port()
    path = str(tmp_path / "result.json")
    mp.spawn(_worker, args=(world, port, path), nprocs=world, join=True)
    with open(path) as fh:
        results = json.load(fh)
    print(json.dumps(results))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "nccl_sharding_test.json"), "w") as fh:
            json.dump(results, fh)
    assert results[0]["period"] == results[1]["period"]
    for r in results:
        assert r["period_bit_equal"] == 1.0 and r["minloc_ok"] == 1.0
        assert r["filter_worst_rel"] <= RTOL
