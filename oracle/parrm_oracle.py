"""CPU oracle for the PARRM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a plain NumPy/SciPy restatement of the reference algorithm
(neuromodulation/PyPARRM, ``src/pyparrm/parrm.py``) for the two hot paths the
B200 build accelerates: the period search behind ``find_period`` and the comb
filter behind ``filter_data``.  It exists so that the CUDA path can be checked
against something that runs anywhere (the reference itself cannot travel to
the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``pyparrm_b200`` never does.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the unmodified reference
(imported from ``/root/reference`` through ``oracle/ref_shim.py``) on seeded
inputs and on the bundled DBS recording, and ``tests/test_oracle_golden.py``
checks every function here against those recorded outputs, including the one
known-answer vector the reference ships (``matlab_filtered.npy``, used by
``examples/plot_use_parrm.py:210-240``).

Third-party arithmetic the reference delegates to (unpinned in its
``pyproject.toml:12``): ``numpy.linalg.solve`` (LAPACK dgesv),
``scipy.optimize.fmin`` (Nelder-Mead) and ``scipy.signal.convolve`` (auto ->
FFT).  The oracle calls the same library entry points, so it is also a fair
stand-in for the reference's CPU cost.

Every function cites the reference lines it restates.
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy.optimize import fmin
from scipy.signal import convolve

SAMPLE_CAPS = (5000, 10000, 25000)  # parrm.py:291
IGNORE_PORTIONS = (0.0, 0.0, 0.95)  # parrm.py:294
BANDWIDTHS = (5, 10, 20)  # parrm.py:295
GRID_LAMBDA = 1.0  # parrm.py:296
N_RESTARTS = 5  # parrm.py:499
HALF_WIDTH_MATCHES = 50  # parrm.py:792


# --------------------------------------------------------------------------
# period search
# --------------------------------------------------------------------------
def standardise(data: np.ndarray, outlier_boundary: float) -> np.ndarray:
    """First difference, scale by mean |diff| per channel, clip (parrm.py:272-280)."""
    with np.errstate(all="ignore"):
        z = np.diff(data, axis=1)
        z /= np.mean(np.abs(z), axis=1)[:, None]
        return np.clip(z, -outlier_boundary, outlier_boundary)


def run_plan(n_search: int) -> list[tuple[int, float, int]]:
    """(use_n, ignore_portion, bandwidth) per search run (parrm.py:288-301).

    The de-duplicated sample caps are *zipped* with the fixed portion and
    bandwidth lists, so short recordings get fewer runs that still start from
    the first list entries.
    """
    lens = np.unique([int(min(n_search, cap)) for cap in SAMPLE_CAPS])
    return [
        (int(n), ig, bw) for n, ig, bw in zip(lens, IGNORE_PORTIONS, BANDWIDTHS)
    ]


def centre_indices(
    search_samples: np.ndarray,
    n_samples: int,
    use_n: int,
    ignore_portion: float,
    rng: np.random.Generator,
) -> np.ndarray:
    """Sample indices fitted in one run (parrm.py:327-374).

    Only the first and last entries of the (sorted) ``search_samples`` are read.
    """
    span = search_samples[0] + search_samples[-1]
    lo = int(np.ceil((span - use_n) / 2))
    hi = int(np.floor((span + use_n) / 2))
    if n_samples * ignore_portion < hi - lo:
        return np.arange(lo, hi + 1)
    lo = int(search_samples[0] + np.floor((1.0 - ignore_portion) / 2.0 * n_samples))
    hi = int(search_samples[-1] - np.ceil((1.0 - ignore_portion) / 2.0 * n_samples))
    draws = rng.integers(0, hi - lo, np.min((use_n, hi - lo)))
    return np.unique(draws) + lo


def candidate_periods(estimates, run: int) -> np.ndarray:
    """Coarse (+-1 %/run) and fine (+-0.1 %/run) candidate grid (parrm.py:376-405)."""
    scale = np.concatenate(
        (
            1 + np.arange(-1e-2, 1e-2 + 1e-4, 1e-4) / run,
            1 + np.arange(-1e-3, 1e-3 + 1e-5, 1e-5) / run,
        )
    )
    out = []
    for est in estimates:
        out.extend(est * scale)
    return np.unique(out)


def harmonic_design(indices: np.ndarray, period, bandwidth: int) -> np.ndarray:
    """Design matrix [N, 2*bw+1] = [1, sin(k a), cos(k a)] (parrm.py:619-623)."""
    angles = (indices + 1) * (2 * np.pi / period)
    waves = np.ones((indices.shape[0], 2 * bandwidth + 1))
    for k in range(1, bandwidth + 1):
        waves[:, 2 * k - 1] = np.sin(k * angles)
        waves[:, 2 * k] = np.cos(k * angles)
    return waves


def fit_harmonics(y: np.ndarray, indices: np.ndarray, period, bandwidth: int):
    """Squared residuals and squared coefficients for one channel (parrm.py:599-632).

    Returns ``(inf, inf)`` when LAPACK reports an exactly singular Gram matrix.
    """
    waves = harmonic_design(indices, period, bandwidth)
    try:
        beta = np.linalg.solve(waves.T @ waves, waves.T @ y)
    except np.linalg.LinAlgError:
        return np.inf, np.inf
    res = y - waves @ beta
    return res**2, beta**2


def objective(
    period,
    z: np.ndarray,
    indices: np.ndarray,
    bandwidth: int,
    lambda_: float,
    n_chans: int | None = None,
) -> float:
    """Fit error of one candidate period, averaged over channels (parrm.py:552-597).

    ``n_chans`` is the divisor the reference takes from the constructor data
    (``self._n_chans``); it defaults to ``z.shape[0]``.
    """
    if n_chans is None:
        n_chans = z.shape[0]
    weights = np.arange(1, 2 * bandwidth + 2)
    weights = lambda_ * weights / weights.sum()
    total = 0.0
    with np.errstate(all="ignore"):
        for chan in z:
            res2, beta2 = fit_harmonics(chan[indices], indices, period, bandwidth)
            if isinstance(res2, float):
                return np.inf
            total += res2.mean() + weights @ beta2
    return total / n_chans


def objective_many(
    periods, z, indices, bandwidth, lambda_, n_chans=None, n_jobs: int = 1
) -> np.ndarray:
    """Map :func:`objective` over candidates on ``n_jobs`` threads.

    Mirrors the ``pqdm.threads`` map of parrm.py:445-454 (order-preserving,
    thread-based; the mapped function is pure so threading cannot change values).
    """
    def one(p):
        return objective(p, z, indices, bandwidth, lambda_, n_chans)

    if n_jobs <= 1:
        return np.array([one(p) for p in periods])
    with ThreadPoolExecutor(max_workers=n_jobs) as pool:
        return np.array(list(pool.map(one, periods)))


def grid_stage(periods, z, indices, bandwidth, lambda_, n_chans=None, n_jobs=1):
    """Evaluate and rank the candidate grid (parrm.py:407-465).

    Returns (periods sorted by error with non-finite ones dropped, all errors
    sorted -- the reference keeps the non-finite tail on the error vector -- and
    the unsorted errors in candidate order).
    """
    raw = objective_many(periods, z, indices, bandwidth, lambda_, n_chans, n_jobs)
    order = raw.argsort()
    err = raw[order]
    ranked = periods[order[np.isfinite(err)]]
    if ranked.shape == (0,):
        raise ValueError(
            "The period cannot be estimated from the data. Check that your data "
            "does not contain infs or NaNs."
        )
    return ranked, err, raw


def refine_stage(ranked, err, z, indices, bandwidth, lambda_, n_chans=None):
    """Nelder-Mead from the best <=5 grid candidates (parrm.py:467-522)."""
    n_iters = int(np.min((N_RESTARTS, ranked.shape[0])))
    for k in range(n_iters):
        out = fmin(
            objective,
            ranked[k],
            (z, indices, bandwidth, lambda_, n_chans),
            full_output=True,
            disp=False,
        )
        ranked[k] = out[0][0]
        err[k] = out[1]
    return (ranked[err.argmin()],)


def final_stage(period, z, indices, bandwidth, n_chans=None):
    """Unregularised Nelder-Mead polish (parrm.py:524-550)."""
    return fmin(objective, period, (z, indices, bandwidth, 0.0, n_chans), disp=False)[0]


def find_period(
    data: np.ndarray,
    sampling_freq: float,
    artefact_freq: float,
    search_samples: np.ndarray | None = None,
    assumed_periods=None,
    outlier_boundary: float = 3.0,
    random_seed: int | None = None,
    n_jobs: int = 1,
    trace: dict | None = None,
):
    """Whole period search (parrm.py:148-194, 282-325).  Returns ``np.float64``.

    ``trace`` (optional dict) receives per-run ``indices``, candidate ``periods``,
    raw grid ``errors`` and the run's refined estimate, for golden comparison.
    """
    n_chans, n_samples = data.shape
    if search_samples is None:
        search_samples = np.arange(n_samples - 1)  # parrm.py:225
    search_samples = np.sort(search_samples)
    if assumed_periods is None:
        assumed_periods = (sampling_freq / artefact_freq,)  # parrm.py:242
    elif isinstance(assumed_periods, (int, float)):
        assumed_periods = (assumed_periods,)

    z = standardise(data, outlier_boundary)
    rng = np.random.default_rng(random_seed)  # parrm.py:284
    estimate = assumed_periods
    indices = None
    for run, (use_n, ignore, bw) in enumerate(run_plan(search_samples.shape[0]), 1):
        indices = centre_indices(search_samples, n_samples, use_n, ignore, rng)
        bw = int(np.min((bw, indices.shape[0] // 4)))  # parrm.py:305
        periods = candidate_periods(estimate, run)
        ranked, err, raw = grid_stage(
            periods, z, indices, bw, GRID_LAMBDA, n_chans, n_jobs
        )
        estimate = refine_stage(ranked, err, z, indices, bw, GRID_LAMBDA, n_chans)
        if trace is not None:
            trace.setdefault("runs", []).append(
                dict(indices=indices, bandwidth=bw, periods=periods, errors=raw,
                     estimate=float(estimate[0]))
            )
    if not np.isfinite(estimate[0]):  # parrm.py:317-321
        raise ValueError(
            "The period cannot be estimated from the data. Check that your data "
            "does not contain infs or NaNs."
        )
    # final run: unclamped last bandwidth, last run's indices (parrm.py:323-325)
    return final_stage(estimate[0], z, indices, BANDWIDTHS[-1], n_chans)


# --------------------------------------------------------------------------
# filter
# --------------------------------------------------------------------------
def default_half_width(
    period: float, period_half_width: float, omit_n_samples: int, n_samples: int
) -> int:
    """Smallest half-width holding 50 phase matches (parrm.py:788-801).

    The second clause is ``>= period + phw`` in the reference (never true); it is
    kept as written.
    """
    hw = omit_n_samples
    hits = 0
    while hits < HALF_WIDTH_MATCHES and hw < (n_samples - 1) // 2:
        hw += 1
        m = np.mod(hw, period)
        if m <= period_half_width or m >= period + period_half_width:
            hits += 1
    return hw


def tap_mask(
    period: float,
    period_half_width: float,
    filter_half_width: int,
    omit_n_samples: int,
    direction: str,
) -> np.ndarray:
    """Boolean mask over window offsets -hw..hw (parrm.py:805-820)."""
    window = np.arange(-filter_half_width, filter_half_width + 1)
    phase = np.mod(window, period)
    keep = (
        (phase <= period_half_width) | (phase >= period - period_half_width)
    ) & (np.abs(window) > omit_n_samples)
    if direction == "past":
        keep[window > 0] = False
    elif direction == "future":
        keep[window <= 0] = False
    return keep


def tap_offsets(period, period_half_width, filter_half_width, omit_n_samples, direction):
    """Signed window offsets ``w`` of the taps, ascending, int32."""
    keep = tap_mask(period, period_half_width, filter_half_width, omit_n_samples, direction)
    return (np.nonzero(keep)[0] - filter_half_width).astype(np.int32)


def build_filter(period, period_half_width, filter_half_width, omit_n_samples, direction):
    """Filter vector: -1/n_taps on taps, 1 at the centre (parrm.py:803-833)."""
    keep = tap_mask(period, period_half_width, filter_half_width, omit_n_samples, direction)
    filt = keep.astype(np.float64)
    if not keep.any():
        raise RuntimeError(
            "A suitable filter cannot be created with the specified settings. Try "
            "reducing the number of omitted samples and/or increasing the filter "
            "half-width."
        )
    filt = -filt / np.max((filt.sum(), np.finfo(np.float64).eps))
    filt[filter_half_width] = 1
    return filt


def apply_filter_fft(data: np.ndarray, filt: np.ndarray) -> np.ndarray:
    """The reference's own arithmetic: two SciPy convolutions (parrm.py:861-869).

    This is the CPU-baseline path.  At edge samples with no tap in range SciPy's
    FFT method returns rounding noise over rounding noise.
    """
    with np.errstate(all="ignore"):
        num = convolve(data.T, filt[:, np.newaxis], "same") - data.T
        den = 1 - convolve(np.ones_like(data).T, filt[:, np.newaxis], "same")
        out = (num / den + data.T).T
        out[~np.isfinite(out)] = 0
    return out


def in_range_tap_count(n_times: int, taps: np.ndarray) -> np.ndarray:
    """n_in(t) = #{w in taps : 0 <= t - w < T} for every t."""
    count = np.zeros(n_times + 1, dtype=np.int64)
    for w in taps.astype(np.int64):
        lo, hi = max(0, w), min(n_times, n_times + w)  # t - w in [0, T)
        if lo < hi:
            count[lo] += 1
            count[hi] -= 1
    return np.cumsum(count[:-1])


def apply_filter_direct(data: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """What parrm.py:861-869 means in exact arithmetic.

    ``y[t] = x[t] - mean{x[t-w] : w in taps, 0 <= t-w < T}``, and 0 where no tap
    is in range (the intent of the non-finite -> 0 rule, parrm.py:867-869).
    Equal to :func:`apply_filter_fft` to ~1e-15 wherever ``n_in(t) > 0``.
    """
    n_chans, n_times = data.shape
    data = data.astype(np.float64, copy=False)
    acc = np.zeros((n_chans, n_times), dtype=np.float64)
    for w in taps.astype(np.int64):
        lo, hi = max(0, w), min(n_times, n_times + w)
        if lo < hi:
            acc[:, lo:hi] += data[:, lo - w : hi - w]
    n_in = in_range_tap_count(n_times, taps)
    out = np.zeros_like(acc)
    ok = n_in > 0
    out[:, ok] = data[:, ok] - acc[:, ok] / n_in[ok]
    return out


def periodogram(data, sampling_freq, n_points, max_freq=None):
    """``compute_psd`` (src/pyparrm/_utils/_power.py:10-68), restated with ``numpy.fft`` in
    single precision steps: the first ``n_points`` samples (cropped / zero-padded), bins
    ``1 .. n_points // 2``, ``float32(|X|)**2 / (fs * n)``, then ``psd[:-1] *= 2`` -- on a
    2-D array that doubles every row but the last (lines 63-66)."""
    n_points = int(n_points)
    freqs = np.abs(np.fft.fftfreq(n_points, 1.0 / sampling_freq)[1:(n_points // 2) + 1])
    if max_freq is None:
        max_freq = freqs[-1]
    last = np.argwhere(freqs <= max_freq)[-1][0]
    x = np.asarray(data).astype(np.float32)
    coeffs = np.fft.fft(x, n_points)[..., 1:(n_points // 2) + 1]
    psd = (1.0 / (sampling_freq * n_points)) * np.abs(coeffs).astype(np.float32) ** 2
    psd = psd.astype(np.float32)
    psd[:-1] *= 2
    return freqs[: last + 1], psd[..., : last + 1]
