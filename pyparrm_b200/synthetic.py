"""Seeded synthetic recordings with an injected periodic stimulation artefact.

Workload generator for the parity tests and ``bench.py``:
unit-variance white "neural" background per channel plus one stimulation
waveform shared by all channels, ``A_c * sum_k a_k sin(2 pi k t / p_true + phi_k)``
with ``p_true = fs / fa * (1 + drift)``, ``a_k = 1/k``, ``A_c ~ U[2, 5]``,
``phi_k ~ U[0, 2 pi)``.
"""

from __future__ import annotations

import numpy as np


def true_period(sampling_freq: float, artefact_freq: float, drift: float = 3e-6) -> float:
    return sampling_freq / artefact_freq * (1.0 + drift)


def make_recording(
    n_chans: int,
    n_samples: int,
    sampling_freq: float,
    artefact_freq: float,
    seed: int = 0,
    n_harmonics: int = 5,
    drift: float = 3e-6,
    dtype=np.float64,
    out: np.ndarray | None = None,
) -> np.ndarray:
    """Return ``[n_chans, n_samples]`` C-contiguous data (written into ``out`` if given)."""
    rng = np.random.default_rng(seed)
    period = true_period(sampling_freq, artefact_freq, drift)
    amps = rng.uniform(2.0, 5.0, n_chans)
    phases = rng.uniform(0.0, 2.0 * np.pi, n_harmonics)
    t = np.arange(n_samples, dtype=np.float64)
    wave = np.zeros(n_samples)
    for k in range(1, n_harmonics + 1):
        wave += np.sin((2.0 * np.pi * k / period) * t + phases[k - 1]) / k
    if out is None:
        out = np.empty((n_chans, n_samples), dtype=dtype)
    for c in range(n_chans):
        out[c] = rng.standard_normal(n_samples) + amps[c] * wave
    return out
