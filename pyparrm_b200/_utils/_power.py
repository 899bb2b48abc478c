"""Signal power: ``compute_psd`` of the reference (``src/pyparrm/_utils/_power.py:10-68``) on
the device.

Same signature, same return types (``freqs`` float64, ``psd`` float32), same quirks: the
spectrum is the periodogram of the FIRST ``n_points`` samples (``scipy.fft.fft(x, n_points)``
crops or zero-pads), the zero frequency is dropped, and ``psd[:-1] *= 2`` doubles everything
but the last ROW of a 2-D input (but the last BIN of a 1-D input).  ``data`` may also be a
CUDA tensor (e.g. the device-resident result of ``DeviceEngine.filter_device``), in which case
nothing but the small spectrum crosses PCIe -- the interactive explorer's
filter -> spectrum loop (``_utils/_plotting.py:568-584, 637-642``) stays on the GPU.
"""

from __future__ import annotations

import numpy as np

from .. import _engine


def compute_psd(data, sampling_freq, n_points: int, max_freq=None, n_jobs: int = 1):
    """Power spectral density of ``data`` ([channels, times] or [times]); see the module
    docstring.  ``n_jobs`` is accepted for compatibility and ignored."""
    n_points = int(n_points)
    n_bins = n_points // 2
    # reference lines 58-61, restated (fftfreq(n, 1/fs)[1 : n//2 + 1] in absolute value)
    freqs = np.abs(np.fft.fftfreq(n_points, 1.0 / sampling_freq)[1: n_bins + 1])
    if max_freq is None:
        max_freq = freqs[-1]
    max_freq_i = np.argwhere(freqs <= max_freq)[-1][0]
    psd = _engine.get_engine().periodogram(data, n_points, float(sampling_freq))
    psd[:-1] *= 2
    return freqs[: max_freq_i + 1], psd[..., : max_freq_i + 1]
