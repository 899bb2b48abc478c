"""The lock-step Nelder-Mead must retrace scipy.optimize.fmin (reference parrm.py:499-517, 545-550)."""

import numpy as np
import pytest
from scipy.optimize import fmin

from pyparrm_b200._neldermead import fmin_batch


def rugged(x):
    x = float(np.asarray(x).ravel()[0])
    return (x - 15.3846) ** 2 * 40 + 0.3 * np.sin(900 * x) + 0.05 * np.cos(13000 * x)


def flat_then_cliff(x):
    x = float(np.asarray(x).ravel()[0])
    return 1.0 if x < 2.0 else 1.0 + 1e-6 * (x - 2.0)


def nan_patch(x):
    x = float(np.asarray(x).ravel()[0])
    return np.nan if 1.36 < x < 1.37 else (x - 1.3311) ** 2


def inf_everywhere(x):
    return np.inf


def never_settles(x):
    x = float(np.asarray(x).ravel()[0])
    return np.sin(1e6 * x) * 1e3


CASES = [
    (rugged, [15.2, 15.3846, 15.39, 15.5, 16.0]),
    (flat_then_cliff, [1.9, 2.0, 2.5]),
    (nan_patch, [1.30, 1.3333333333, 1.40]),
    (inf_everywhere, [1.3333, 7.7]),
    (never_settles, [2.0, 3.0]),
    (lambda x: float(np.asarray(x).ravel()[0]) ** 2, [0.0, 1e-9, -3.0]),
]


@pytest.mark.parametrize("func,starts", CASES)
def test_matches_scipy_fmin(func, starts):
    def batch(points):
        return np.array([func(p) for p in points], dtype=np.float64)

    mine = fmin_batch(batch, starts)
    for x0, (x, fval, nit, nfev) in zip(starts, mine):
        ref = fmin(func, x0, full_output=True, disp=False)
        assert np.array_equal(np.float64(x), ref[0][0], equal_nan=True)
        assert np.array_equal(np.float64(fval), np.float64(ref[1]), equal_nan=True)
        assert (nit, nfev) == (ref[2], ref[3])


def test_batch_rounds_are_shared():
    rounds = []

    def batch(points):
        rounds.append(len(points))
        return np.array([rugged(p) for p in points])

    fmin_batch(batch, [15.2, 15.4, 15.5])
    assert rounds[0] == 6 and all(r % 5 == 0 for r in rounds[1:])
    assert len(rounds) < 80  # one launch per iteration, not one per evaluation
