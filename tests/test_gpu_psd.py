"""Device compute_psd vs the reference function's recorded outputs (tests/golden/psd.npz,
made by oracle/make_golden.py from src/pyparrm/_utils/_power.py).  float32 result: agreement
to 1e-4 of the largest bin (BASELINE north_star fp32 tolerance); frequencies bit-equal."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_compute_psd_matches_reference(golden, gpu_engine):
    from pyparrm_b200._utils._power import compute_psd

    g = golden("psd")
    for k in range(int(g["n_cases"])):
        fs, n, fmax = g[f"case{k}_args"]
        x = g[f"case{k}_x"]
        freqs, psd = compute_psd(data=x, sampling_freq=fs, n_points=int(n),
                                 max_freq=None if fmax < 0 else fmax)
        want = g[f"case{k}_psd"]
        assert isinstance(freqs, np.ndarray) and isinstance(psd, np.ndarray)
        assert np.array_equal(freqs, g[f"case{k}_freqs"])
        assert psd.dtype == np.float32 and psd.shape == want.shape, k
        assert np.abs(psd - want).max() <= 1e-4 * np.abs(want).max(), k


def test_reference_suite_shapes(gpu_engine):
    """tests/test_utils.py::test_compute_psd of the reference, against this build."""
    from pyparrm_b200._utils._power import compute_psd

    data = np.random.default_rng(44).standard_normal((2, 100))
    freqs, psd = compute_psd(data=data, sampling_freq=20, n_points=10, n_jobs=2)
    assert psd.shape == (2, 5) and freqs.shape[0] == psd.shape[1]
    max_freq = (20 / 2) - ((20 / 2) / 5)
    freqs, psd = compute_psd(data=data, sampling_freq=20, n_points=10, max_freq=max_freq)
    assert freqs.shape[0] == psd.shape[1] and freqs[-1] == max_freq


def test_filter_then_spectrum_stays_on_device(gpu_engine):
    """The explorer's loop (_plotting.py:568-584, 637-642): filter, then the spectrum of one
    channel -- here from the device-resident result, equal to the host route."""
    import torch

    from oracle import parrm_oracle as oracle
    from pyparrm_b200._utils._power import compute_psd
    from pyparrm_b200.synthetic import make_recording

    x = make_recording(3, 60_000, 2000, 130, seed=4)
    taps = oracle.tap_offsets(2000 / 130, 0.3, 2000, 0, "both")
    d_y = gpu_engine.filter_device(torch.from_numpy(x).cuda(), taps)
    f_dev, p_dev = compute_psd(d_y[1], 2000, 400, max_freq=500.0)
    f_host, p_host = compute_psd(d_y[1].cpu().numpy(), 2000, 400, max_freq=500.0)
    assert np.array_equal(f_dev, f_host) and np.array_equal(p_dev, p_host)
    _, want = oracle.periodogram(oracle.apply_filter_direct(x, taps)[1], 2000, 400, 500.0)
    assert np.abs(p_dev - want).max() <= 1e-4 * np.abs(want).max()
    from pyparrm_b200 import install_as_pyparrm

    install_as_pyparrm()
    from pyparrm._utils._power import compute_psd as aliased

    assert aliased is compute_psd
