"""Filter planner (host side of the C ABI) and the strip-kernel design, on the CPU.

The planner regroups the tap set of ``_generate_filter`` (parrm.py:803-833) into comb boxes
plus single taps (pyparrm_b200/csrc/filter_plan.h).  That is only legal if it is an identity
over integers, so every plan is expanded back and compared with the tap set, bit for bit.
"""

import numpy as np
import pytest

from oracle import parrm_oracle as oracle
from pyparrm_b200 import _native
from tests.strip_model import StripModel


def expand(desc, lo, hi):
    acc = np.zeros(hi - lo + 1, dtype=np.int64)
    for m, boxes in zip(desc["windows"], desc["boxes"]):
        for a in boxes:
            for q in range(m):
                acc[a + q * desc["stride"] - lo] += 1
    for w in desc["plus"]:
        acc[w - lo] += 1
    for w in desc["minus"]:
        acc[w - lo] -= 1
    acc[0 - lo] += desc["centre"]
    return acc


CASES = [
    # period, phw, hw, omit, direction   (BASELINE configs first)
    (2000 / 130 * (1 + 3e-6), None, 2000, 0, "both"),
    (1000 / 145 * (1 + 3e-6), None, 2469, 0, "both"),
    (30000 / 130 * (1 + 3e-6), None, 2311, 0, "past"),
    (30000 / 130 * (1 + 3e-6), None, 2311, 0, "future"),
    (1.3311148014466094, 0.01, 2000, 20, "both"),
    (1.3311148014466094, None, 2000, 0, "both"),
    (15.3846, None, 777, 3, "past"),
    (15.3846, None, 50, 0, "future"),
    (2.0, None, 40, 0, "both"),
    (7.123456, 0.9, 3000, 100, "both"),
    (101.5, 3.3, 5000, 0, "both"),
    (230.77, 4.6, 40_000, 0, "both"),
    (2000 / 130 * (1 + 3e-6), None, 1500, 300, "both"),
]


@pytest.mark.parametrize("case", CASES)
def test_plans_are_exact_regroupings(case):
    period, phw, hw, omit, direction = case
    taps = oracle.tap_offsets(period, period / 50 if phw is None else phw, hw, omit, direction)
    for strategy in (_native.PLAN_AUTO, _native.PLAN_GATHER):
        _, desc = _native.plan_filter(taps, strategy=strategy)
        assert desc["n_taps"] == len(taps)
        if strategy == _native.PLAN_GATHER:
            assert desc["kind"] == 0
        if desc["kind"] == 0:
            continue
        lo, hi = min(int(taps[0]), 0), max(int(taps[-1]), 0)
        want = np.zeros(hi - lo + 1, dtype=np.int64)
        want[taps - lo] = 1
        assert np.array_equal(expand(desc, lo, hi), want)
        n_terms = sum(len(b) for b in desc["boxes"]) + len(desc["plus"]) + len(desc["minus"])
        assert n_terms <= 120 and desc["cost"] < 0.6 * len(taps)
        assert desc["centre"] <= 0 and len(desc["windows"]) in (1, 2)


def test_baseline_configs_get_short_plans():
    """cfg2 (160 taps) must cost ~10 loads, not 160: that is what the roofline target needs."""
    period = 2000 / 130 * (1 + 3e-6)
    taps = oracle.tap_offsets(period, period / 50, 2000, 0, "both")
    _, desc = _native.plan_filter(taps)
    assert desc["kind"] == 1 and desc["stride"] == 200 and sorted(desc["windows"]) == [10, 20]
    assert sum(len(b) for b in desc["boxes"]) + len(desc["plus"]) + len(desc["minus"]) <= 12


def test_random_tap_sets():
    rng = np.random.default_rng(7)
    for _ in range(40):
        n = int(rng.integers(1, 300))
        taps = np.unique(rng.integers(-3000, 3000, n))
        taps = taps[taps != 0].astype(np.int32)
        if len(taps) == 0:
            continue
        _, desc = _native.plan_filter(taps)
        if desc["kind"] == 1:
            lo, hi = min(int(taps[0]), 0), max(int(taps[-1]), 0)
            want = np.zeros(hi - lo + 1, dtype=np.int64)
            want[taps - lo] = 1
            assert np.array_equal(expand(desc, lo, hi), want)
    one, blob = np.array([5], dtype=np.int32), np.zeros(4096, dtype=np.uint8)
    status = _native.lib.parrm_filter_plan(one.ctypes.data, 1, 0, _native.PLAN_COMB,
                                           blob.ctypes.data, 4096)
    assert status == 4 and "comb" in _native.last_error()


MODEL_CASES = [
    # case index, recording length, tile, prefetch, gamma, pieces, reinit, time chunk
    (0, 30_011, 512, 3, 1, 3, 0, None),
    (0, 700, 512, 3, 0, 1, 0, None),
    (0, 20_000, 256, 4, 0, 2, 5, None),
    (0, 40_000, 1024, 2, 1, 2, 0, (9_000, 21_000)),
    (2, 20_000, 512, 3, 0, 2, 0, None),
    (3, 20_000, 1024, 2, 1, 1, 7, None),
    (4, 19_130, 512, 3, 0, 1, 0, None),
    (5, 19_130, 512, 3, 0, 4, 0, None),
    (12, 15_000, 256, 2, 0, 3, 0, (0, 7_000)),
]


@pytest.mark.parametrize("pipe", [0, 2, 3])
@pytest.mark.parametrize("spec", MODEL_CASES)
def test_strip_design_matches_oracle(spec, pipe):
    """Ring slots, mirror chunk, sliding boxes, piece boundaries and edge counts of the strip
    kernels (NumPy model with their index arithmetic) against the oracle's direct sum.
    ``pipe`` = number of hand-over stages of the producer/consumer kernel (0: two-phase kernel);
    the model runs the slide the full ``stages - 1`` chunks ahead of the gather."""
    case, n_total, tile, prefetch, gamma, pieces, reinit, chunk = spec
    period, phw, hw, omit, direction = CASES[case]
    taps = oracle.tap_offsets(period, period / 50 if phw is None else phw, hw, omit, direction)
    _, desc = _native.plan_filter(taps, strategy=_native.PLAN_COMB)
    rng = np.random.default_rng(case)
    x = rng.standard_normal((1, n_total)) + 3.0
    want = oracle.apply_filter_direct(x, taps)[0]
    model = StripModel(taps, desc, tile, prefetch, 0 if pipe else reinit, pipe=bool(pipe),
                       stages=max(pipe, 2))
    if chunk is None:
        got = model.run(x[0], 0, 0, n_total, n_total, gamma, pieces)
    else:
        t0, t1 = chunk
        x0, x1 = max(0, t0 - model.w_hi), min(n_total, t1 - model.w_lo)
        got = model.run(x[0, x0:x1], x0, t0, t1 - t0, n_total, gamma, pieces)
        want = want[t0:t1]
    assert not np.isnan(got).any()
    assert np.abs(got - want).max() <= 1e-12
