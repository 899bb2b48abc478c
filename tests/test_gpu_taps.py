"""Device tap builder == reference's NumPy mask, bit for bit (parrm.py:803-820)."""

import numpy as np
import pytest

from oracle import parrm_oracle as oracle

pytestmark = pytest.mark.gpu
DIRECTIONS = ("both", "past", "future")


def test_golden_parameter_sweep(golden, gpu_engine):
    g = golden("taps")
    table, starts, taps = g["table"], g["starts"], g["taps"]
    for row, (per, phw, hw, omit, d, n, n_taps) in enumerate(table):
        mine = gpu_engine.build_taps(per, phw, int(hw), int(omit), DIRECTIONS[int(d)])
        assert mine.dtype == np.int32
        assert np.array_equal(mine, taps[starts[row] : starts[row + 1]]), table[row]
        assert (mine.shape[0] == 0) == (n_taps < 0)


def test_random_parameters_against_oracle(gpu_engine):
    rng = np.random.default_rng(5)
    for _ in range(200):
        per = float(rng.uniform(0.6, 400.0)) if rng.random() < 0.7 else float(rng.integers(1, 50)) / float(rng.integers(1, 9))
        phw = float(rng.uniform(1e-4, 1.0) * per) if rng.random() < 0.8 else per / 50
        hw = int(rng.integers(1, 6000))
        omit = int(rng.integers(0, max(1, hw // 3)))
        d = DIRECTIONS[int(rng.integers(0, 3))]
        assert np.array_equal(gpu_engine.build_taps(per, phw, hw, omit, d),
                              oracle.tap_offsets(per, phw, hw, omit, d)), (per, phw, hw, omit, d)


def test_wide_window(gpu_engine):
    per = 3000 / 13 * (1 + 3e-6)
    assert np.array_equal(gpu_engine.build_taps(per, per / 50, 300_000, 11, "both"),
                          oracle.tap_offsets(per, per / 50, 300_000, 11, "both"))


def test_whole_golden_sweep_in_one_launch(golden, gpu_engine):
    """All 288 parameter sets of the golden sweep (tests/golden/taps.npz, recorded from the
    reference's _generate_filter) built by ONE parrm_build_taps_batch launch: bit-exact."""
    import torch

    from pyparrm_b200 import _native
    from pyparrm_b200._engine import _vp

    g = golden("taps")
    table, starts, taps = g["table"], g["starts"], g["taps"]
    n = table.shape[0]
    stride = int(2 * table[:, 2].max() + 1)
    dev = lambda a, dt: torch.tensor(np.ascontiguousarray(a, dtype=dt), device="cuda")  # noqa: E731
    d_per, d_phw = dev(table[:, 0], np.float64), dev(table[:, 1], np.float64)
    d_hw, d_omit = dev(table[:, 2], np.int64), dev(table[:, 3], np.int64)
    d_dir = dev(table[:, 4], np.int32)
    d_taps = torch.empty((n, stride), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(n, dtype=torch.int32, device="cuda")
    launches0 = gpu_engine.launches
    _native.check(_native.lib.parrm_build_taps_batch(
        _vp(d_per.data_ptr()), _vp(d_phw.data_ptr()), _vp(d_hw.data_ptr()), _vp(d_omit.data_ptr()),
        _vp(d_dir.data_ptr()), n, _vp(d_taps.data_ptr()), stride, _vp(d_n.data_ptr()),
        _vp(torch.cuda.current_stream().cuda_stream)), "parrm_build_taps_batch")
    counts, rows = d_n.cpu().numpy(), d_taps.cpu().numpy()
    for row in range(n):
        assert np.array_equal(rows[row, : counts[row]], taps[starts[row]: starts[row + 1]]), table[row]
    # default half-widths of the same sweep (parrm.py:788-801), one launch, vs the reference's
    want = g["default_half_width"]
    got = gpu_engine.default_half_widths(table[:, 0], table[:, 1], table[:, 3].astype(np.int64),
                                         int((table[0, 5] - 1) // 2))
    same_n = table[:, 5] == table[0, 5]
    assert np.array_equal(got[same_n], want[same_n])
    for per, phw, omit, n_samples in [(15.3846, 0.3, 0, 100), (2.0, 0.04, 3, 31), (7.7, 7.7, 0, 5001),
                                      (101.5, 0.01, 10, 1_200_000)]:
        got = gpu_engine.default_half_widths([per], [phw], [omit], (n_samples - 1) // 2)
        assert int(got[0]) == oracle.default_half_width(per, phw, omit, n_samples)


def test_filter_sweep_matches_one_by_one(gpu_engine):
    """PARRM.filter_sweep (one tap launch + one filter launch for all sets) against
    create_filter + filter_data per set and against the oracle."""
    from pyparrm_b200 import PARRM
    from pyparrm_b200.synthetic import make_recording

    data = make_recording(2, 9_000, 2000, 130, seed=8)
    parrm = PARRM(data, 2000, 130, verbose=False)
    parrm._period = np.float64(2000 / 130 * (1 + 3e-6))
    sets = [dict(filter_half_width=2000), dict(), dict(filter_direction="past", omit_n_samples=5),
            dict(filter_half_width=700, period_half_width=1.1, filter_direction="future"),
            dict(filter_half_width=40, period_half_width=0.001)]   # last one: no taps at all
    launches0 = gpu_engine.launches
    out, taps = parrm.filter_sweep(sets)
    assert gpu_engine.launches - launches0 == 3          # default widths, taps, filter
    assert out.shape == (len(sets), 2, 9_000)
    for k, s in enumerate(sets):
        single = PARRM(data, 2000, 130, verbose=False)
        single._period = parrm._period
        try:
            single.create_filter(**s)
        except RuntimeError:
            assert taps[k].size == 0 and np.all(out[k] == 0)
            continue
        half = single._filter_half_width
        assert np.array_equal(taps[k], np.flatnonzero(single.filter < 0) - half)
        want = oracle.apply_filter_direct(data, taps[k])
        assert np.abs(out[k] - want).max() <= 1e-13 * np.abs(data).max()
        assert np.abs(out[k] - single.filter_data()).max() <= 1e-12 * np.abs(data).max()
    with pytest.raises(ValueError, match="`filter_half_width` must lie in the range"):
        parrm.filter_sweep([dict(filter_half_width=5000)])
