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
