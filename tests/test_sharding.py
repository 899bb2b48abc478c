"""Multi-rank host logic on the CPU: world size 2 over gloo with the oracle-backed engine.

The N > 1 path of the product is: shard the candidate grid over ranks + one all-gather
(find_period), shard channels -- or time with halos -- with no collective (filter_data).
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyparrm_b200 import _sharding


def test_block_partitions_cover_everything():
    for n in (0, 1, 5, 381, 388, 1000):
        for world in (1, 2, 3, 8):
            blocks = [_sharding.block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == blocks[i + 1][0] for i, b in enumerate(blocks[:-1]))


def test_filter_shards_by_channel_then_time():
    shards = _sharding.channel_or_time_shards(64, 1000, 8, -20, 20)
    assert [s[:2] for s in shards] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert all(s[2:] == (0, 1000, 0, 1000) for s in shards)
    shards = _sharding.channel_or_time_shards(2, 1000, 8, -20, 30)   # 4 ranks per channel
    assert [s[0] for s in shards] == [0, 0, 0, 0, 1, 1, 1, 1]
    assert [s[2:4] for s in shards[:4]] == [(0, 250), (250, 500), (500, 750), (750, 1000)]
    assert shards[1][4:] == (220, 520) and shards[0][4:] == (0, 270) and shards[3][4:] == (720, 1000)
    shards = _sharding.channel_or_time_shards(3, 100, 4, -5, 5)      # one rank idle
    assert shards[3] == (0, 0, 0, 0, 0, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import parrm_oracle as oracle
        from pyparrm_b200 import PARRM, _engine, enable_sharding
        from pyparrm_b200.synthetic import make_recording
        from tests.oracle_engine import OracleEngine

        engine = OracleEngine()
        _engine.set_engine(engine)
        data = make_recording(2, 6000, 200, 13, seed=5)
        # unsharded period first, then the sharded one: must be bit-identical on every rank
        p0 = PARRM(data, 200, 13, verbose=False)
        p0.find_period(random_seed=0)
        calls_unsharded = engine.evaluations
        engine.evaluations = 0
        enable_sharding()
        p1 = PARRM(data, 200, 13, verbose=False)
        p1.find_period(random_seed=0)
        calls_sharded = engine.evaluations

        # filter: channel shards, then time shards (1 channel over 2 ranks), no collective
        taps = oracle.tap_offsets(p1.period, p1.period / 50, 300, 0, "both")
        out, (c0, c1, t0, t1), _ = _sharding.filter_sharded(engine, data, taps)
        want = oracle.apply_filter_direct(data, taps)
        chan_ok = np.array_equal(out, want[c0:c1, t0:t1]) and (c1 - c0, t1 - t0) == (1, 6000)
        out1, (c0, c1, t0, t1), _ = _sharding.filter_sharded(engine, data[:1], taps)
        time_ok = np.allclose(out1, want[c0:c1, t0:t1], rtol=0, atol=1e-12) and t1 - t0 == 3000
        # through the public API: stitched result on rank 0 ("rank0"), everywhere ("all"),
        # own shard only ("none"); three channels over two ranks = uneven blocks
        data3 = make_recording(3, 4000, 200, 13, seed=6)
        want3 = oracle.apply_filter_direct(data3, taps)
        for mode in ("rank0", "all", "none"):
            enable_sharding(gather=mode)
            p3 = PARRM(data3, 200, 13, verbose=False)
            p3._period = p1.period
            p3.create_filter(300, 0, "both")
            got = p3.filter_data()
            c0, c1, t0, t1 = p3.filter_shard
            full = mode == "all" or (mode == "rank0" and rank == 0)
            ref = want3 if full else want3[c0:c1, t0:t1]
            chan_ok = chan_ok and got.shape == ref.shape and np.allclose(got, ref, rtol=0, atol=1e-12)
        # one channel over two ranks through the API (time shards + halos), stitched everywhere
        enable_sharding(gather="all")
        p4 = PARRM(data3[:1], 200, 13, verbose=False)
        p4._period = p1.period
        p4.create_filter(300, 0, "both")
        time_ok = time_ok and np.allclose(p4.filter_data(), want3[:1], rtol=0, atol=1e-12)
        # winner of a sweep with one (error, index) pair per rank
        sweep = p1.period * (1 + np.linspace(-1e-3, 1e-3, 31))
        z = oracle.standardise(data, 3.0)
        idx = np.arange(100, 1100)
        tile = engine.tile_from_standardised(z, idx)
        errs = engine.evaluate(tile, sweep, 5, 1.0, 2)
        best = _sharding.minloc_sharded(lambda blk: engine.evaluate(tile, blk, 5, 1.0, 2), sweep)
        chan_ok = chan_ok and best == (int(np.argmin(errs)), float(errs.min()))

        gathered = [None] * world
        dist.all_gather_object(gathered, (float(p0.period), float(p1.period), calls_unsharded,
                                          calls_sharded, chan_ok, time_ok))
        if rank == 0:
            np.save(result_path, np.array(gathered, dtype=np.float64))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world_size_two_gloo(tmp_path):
    world, port = 2, _free_port()
    path = str(tmp_path / "result.npy")
    mp.spawn(_worker, args=(world, port, path), nprocs=world, join=True)
    res = np.load(path)
    p_unsharded, p_sharded, calls0, calls1, chan_ok, time_ok = res.T
    assert p_unsharded[0] == p_unsharded[1] == p_sharded[0] == p_sharded[1]
    assert np.all(chan_ok == 1) and np.all(time_ok == 1)
    # each rank evaluated about half of every grid (the Nelder-Mead rounds are replicated)
    assert np.all(calls1 < 0.75 * calls0)


def test_cpulist_parsing():
    from pyparrm_b200._sharding import _parse_cpulist

    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
