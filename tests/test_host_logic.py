"""Host side of pyparrm_b200.PARRM on CPU: API conformance with the reference's own tests
(/root/reference/tests/test_parrm.py) and bit-exact search logic against golden periods.

The arithmetic comes from the oracle-backed stand-in engine (tests/oracle_engine.py); what is
under test is everything the product does in Python around the kernels.
"""

import copy
from multiprocessing import cpu_count

import numpy as np
import pytest

from pyparrm_b200 import PARRM, get_example_data_paths
from pyparrm_b200.synthetic import make_recording

FS, FA = 20, 10  # the reference's test rates: assumed period exactly 2.0 (degenerate fits)


def noise(n_chans, n_samples, seed=44):
    return np.random.default_rng(seed).standard_normal((n_chans, n_samples))


@pytest.mark.parametrize("n_chans", [1, 2])
@pytest.mark.parametrize("n_samples", [100, 300])
@pytest.mark.parametrize("search_portion", [None, 0.5])
@pytest.mark.parametrize("n_jobs", [1, -1])
def test_workflow_runs_like_reference_suite(cpu_engine, capsys, n_chans, n_samples,
                                            search_portion, n_jobs):
    """Mirror of test_parrm.py:17-68 (small sizes; the 25 000-sample cases run on the GPU)."""
    data = noise(n_chans, n_samples)
    verbose = n_jobs == 1
    parrm = PARRM(data=data, sampling_freq=FS, artefact_freq=FA, verbose=verbose)
    search = None if search_portion is None else np.arange(0, n_samples * search_portion)
    parrm.find_period(search_samples=search, assumed_periods=FS / FA, random_seed=44, n_jobs=n_jobs)
    for direction in ["future", "past", "both"]:
        parrm.create_filter(filter_direction=direction)
    filtered = parrm.filter_data()
    assert filtered.shape == data.shape and isinstance(filtered, np.ndarray)
    other = noise(1, 50, seed=3)
    assert parrm.filter_data(other).shape == other.shape
    assert repr(parrm) == (
        f"PARRM object | Data: ({n_chans} channels x {n_samples} times) | "
        f"Period: {parrm.period :.4f}"
    )
    if n_jobs == -1:
        assert parrm._n_jobs == cpu_count()
    printed = capsys.readouterr().out
    assert ("Finding the artefact period..." in printed) == verbose


def test_attributes_mirror_private_state(cpu_engine):
    """test_parrm.py:71-109."""
    data = noise(1, 100)
    parrm = PARRM(data=data, sampling_freq=FS, artefact_freq=FA, verbose=False)
    parrm.find_period()
    parrm.create_filter()
    filtered = parrm.filter_data()
    assert filtered is parrm._filtered_data and filtered is parrm.filtered_data
    assert parrm.data is data and parrm._data is data
    assert parrm._period == parrm.period and isinstance(parrm.period, np.float64)
    assert np.all(parrm._filter == parrm.filter)
    s = parrm.settings
    assert s["data"] == {"sampling_freq": FS, "artefact_freq": FA}
    assert np.all(s["period"]["search_samples"] == parrm._search_samples)
    assert s["period"]["assumed_periods"] == parrm._assumed_periods == (FS / FA,)
    assert s["period"]["outlier_boundary"] == parrm._outlier_boundary == 3.0
    assert s["period"]["random_seed"] == parrm._random_seed
    assert s["filter"]["filter_half_width"] == parrm._filter_half_width
    assert s["filter"]["omit_n_samples"] == parrm._omit_n_samples == 0
    assert s["filter"]["filter_direction"] == parrm._filter_direction == "both"
    assert s["filter"]["period_half_width"] == parrm._period_half_width == parrm.period / 50
    assert parrm._standard_data.shape == (1, 99)
    clone = copy.deepcopy(parrm)  # the reference's explorer deep-copies the object
    assert clone.period == parrm.period and clone._data is not parrm._data


def test_type_errors(cpu_engine):
    """test_parrm.py:112-197: same exception types, same messages."""
    data = noise(1, 100)
    with pytest.raises(TypeError, match="`data` must be a NumPy array."):
        PARRM(data=data.tolist(), sampling_freq=FS, artefact_freq=FA)
    with pytest.raises(TypeError, match="`sampling_freq` must be an int or a float."):
        PARRM(data=data, sampling_freq=[FS], artefact_freq=FA)
    with pytest.raises(TypeError, match="`artefact_freq` must be an int or a float."):
        PARRM(data=data, sampling_freq=FS, artefact_freq=[FA])
    with pytest.raises(TypeError, match="`verbose` must be a bool."):
        PARRM(data=data, sampling_freq=FS, artefact_freq=FA, verbose="no")
    parrm = PARRM(data=data, sampling_freq=FS, artefact_freq=FA, verbose=False)
    with pytest.raises(TypeError, match="`search_samples` must be a NumPy array or None."):
        parrm.find_period(search_samples=0)
    with pytest.raises(TypeError, match="`assumed_periods` must be an int, a float, a tuple, or None."):
        parrm.find_period(assumed_periods=[0])
    with pytest.raises(TypeError, match="If a tuple, entries of `assumed_periods` must be ints or floats."):
        parrm.find_period(assumed_periods=(None,))
    with pytest.raises(TypeError, match="`outlier_boundary` must be an int or a float."):
        parrm.find_period(outlier_boundary=[0])
    with pytest.raises(TypeError, match="`random_seed` must be an int or None."):
        parrm.find_period(random_seed=1.5)
    with pytest.raises(TypeError, match="`n_jobs` must be an int."):
        parrm.find_period(n_jobs=1.5)
    parrm.find_period()
    with pytest.raises(TypeError, match="`filter_half_width` must be an int."):
        parrm.create_filter(filter_half_width=1.5)
    with pytest.raises(TypeError, match="`omit_n_samples` must be an int."):
        parrm.create_filter(omit_n_samples=1.5)
    with pytest.raises(TypeError, match="`filter_direction` must be a str."):
        parrm.create_filter(filter_direction=0)
    with pytest.raises(TypeError, match="`period_half_width` must be an int or a float."):
        parrm.create_filter(period_half_width=[0])
    parrm.create_filter()
    with pytest.raises(TypeError, match="`data` must be a NumPy array."):
        parrm.filter_data(data=data.tolist())


def test_value_errors(cpu_engine):
    """test_parrm.py:200-299."""
    data = noise(1, 100)
    with pytest.raises(ValueError, match="`data` must be a 2D array."):
        PARRM(data=noise(1, 100).ravel(), sampling_freq=FS, artefact_freq=FA)
    with pytest.raises(ValueError, match="`sampling_freq` must be > 0."):
        PARRM(data=data, sampling_freq=0, artefact_freq=FA)
    with pytest.raises(ValueError, match="`artefact_freq` must be > 0."):
        PARRM(data=data, sampling_freq=FS, artefact_freq=0)
    with pytest.raises(ValueError, match="`precision` must be"):
        PARRM(data=data, sampling_freq=FS, artefact_freq=FA, precision="bf16")
    parrm = PARRM(data=data, sampling_freq=FS, artefact_freq=FA, verbose=False)
    with pytest.raises(ValueError, match="`search_samples` must be a 1D array."):
        parrm.find_period(search_samples=np.zeros((1, 1)))
    with pytest.raises(ValueError, match=r"Entries of `search_samples` must lie in the range \[0, n_samples\)."):
        parrm.find_period(search_samples=np.array([-1, 1]))
    with pytest.raises(ValueError, match=r"Entries of `search_samples` must lie in the range \[0, n_samples\)."):
        parrm.find_period(search_samples=np.array([0, 100]))
    with pytest.raises(ValueError, match="`outlier_boundary` must be > 0."):
        parrm.find_period(outlier_boundary=0)
    with pytest.raises(ValueError, match="`n_jobs` must be <= the number of available CPUs."):
        parrm.find_period(n_jobs=cpu_count() + 1)
    with pytest.raises(ValueError, match="If `n_jobs` is <= 0, it must be -1."):
        parrm.find_period(n_jobs=-2)
    parrm.find_period()
    half = (100 - 1) // 2
    with pytest.raises(ValueError, match=r"`filter_half_width` must lie in the range"):
        parrm.create_filter(filter_half_width=0)
    with pytest.raises(ValueError, match=r"`filter_half_width` must lie in the range"):
        parrm.create_filter(filter_half_width=half + 1)
    with pytest.raises(ValueError, match=r"`omit_n_samples` must lie in the range"):
        parrm.create_filter(omit_n_samples=-1)
    with pytest.raises(ValueError, match=r"`omit_n_samples` must lie in the range"):
        parrm.create_filter(omit_n_samples=half)
    with pytest.raises(ValueError, match="`filter_direction` must be one of"):
        parrm.create_filter(filter_direction="sideways")
    with pytest.raises(ValueError, match=r"`period_half_width` must be lie in the range \(0, period\]."):
        parrm.create_filter(period_half_width=0)
    with pytest.raises(ValueError, match=r"`period_half_width` must be lie in the range \(0, period\]."):
        parrm.create_filter(period_half_width=parrm.period + 1)
    with pytest.raises(RuntimeError, match="A suitable filter cannot be created with the specified settings."):
        parrm.create_filter(omit_n_samples=48)
    parrm.create_filter()
    with pytest.raises(ValueError, match="`data` must be a 2D array."):
        parrm.filter_data(data=noise(1, 100).ravel())


def test_premature_calls_and_defaults(cpu_engine):
    """test_parrm.py:302-346."""
    parrm = PARRM(data=noise(1, 100), sampling_freq=FS, artefact_freq=FA, verbose=False)
    with pytest.raises(ValueError, match="The period has not yet been estimated."):
        parrm.create_filter()
    with pytest.raises(ValueError, match="The period has not yet been estimated."):
        parrm.explore_filter_params()
    with pytest.raises(ValueError, match="The filter has not yet been created."):
        parrm.filter_data()
    with pytest.raises(AttributeError, match="No period has been computed yet."):
        parrm.period
    with pytest.raises(AttributeError, match="No filter has been computed yet."):
        parrm.filter
    with pytest.raises(AttributeError, match="No data has been filtered yet."):
        parrm.filtered_data
    with pytest.raises(AttributeError, match="Analysis settings have not been established yet."):
        parrm.settings
    parrm.find_period()
    parrm.create_filter()
    assert parrm._filter_half_width is not None and parrm._period_half_width is not None
    parrm.find_period()  # results are reset when the period is re-estimated
    with pytest.raises(AttributeError):
        parrm.filter


def test_default_half_width_matches_reference(golden, cpu_engine):
    g = golden("taps")
    parrm = PARRM(np.zeros((1, 8)), 1, 1, verbose=False)
    for row, (per, phw, hw, omit, d, n, n_taps) in enumerate(g["table"]):
        parrm._n_samples, parrm._period = int(n), np.float64(per)
        parrm._period_half_width, parrm._omit_n_samples = float(phw), int(omit)
        assert parrm._get_filter_half_width() == g["default_half_width"][row], g["table"][row]


def test_candidate_grid_and_indices_match_reference(golden, cpu_engine):
    for name, fs_fa in (("example_dbs", None), ("synthetic_2x30000", None), ("ecog_lfp", None)):
        g = golden(name)
        if name == "synthetic_2x30000":
            n_chans, n, fs, fa, seed = (int(v) for v in g["recording"])
            data, seed_arg = np.zeros((n_chans, n)), 0
        elif name == "ecog_lfp":
            data, seed_arg = np.zeros((2, 60001)), 0
        else:
            data, seed_arg = np.zeros((1, 19130)), None
        parrm = PARRM(data, 1000, 130, verbose=False)
        parrm._search_samples = np.arange(data.shape[1] - 1)
        rng = np.random.default_rng(seed_arg)
        lens = np.unique([min(data.shape[1] - 1, cap) for cap in (5000, 10000, 25000)])
        for run, (use_n, ignore) in enumerate(zip(lens, (0.0, 0.0, 0.95))):
            idx = parrm._get_centre_indices(use_n, ignore, rng)
            np.testing.assert_array_equal(idx, g[f"run{run}_indices"])
            grid = parrm._get_possible_periods(tuple(g[f"run{run}_estimate"]), run + 1)
            np.testing.assert_array_equal(grid, g[f"run{run}_periods"])


def test_find_period_reproduces_reference_bit_for_bit(golden, cpu_engine):
    """Host search logic + lock-step Nelder-Mead over the oracle objective == reference."""
    g = golden("synthetic_2x30000")
    n_chans, n, fs, fa, seed = (int(v) for v in g["recording"])
    data = make_recording(n_chans, n, fs, fa, seed=seed)
    parrm = PARRM(data, fs, fa, verbose=False)
    parrm.find_period(random_seed=0)
    assert parrm.period == g["period"]
    # every period the reference evaluated was also evaluated here (speculation adds more)
    assert cpu_engine.rounds < 400 < len(g["calls"])
    for direction in ("both", "past", "future"):
        parrm.create_filter(filter_direction=direction)
        assert parrm._filter_half_width == int(g["default_half_width"])
        np.testing.assert_array_equal(
            np.flatnonzero(parrm.filter < 0) - parrm._filter_half_width, g[f"{direction}_taps"])


def test_example_recording_period(golden, cpu_engine):
    g = golden("example_dbs")
    data = np.load(get_example_data_paths("example_data"))
    parrm = PARRM(data, 200, 150, verbose=False)
    parrm.find_period()
    assert parrm.period == g["period"]
    parrm.create_filter(filter_half_width=2000, omit_n_samples=20, filter_direction="both",
                        period_half_width=0.01)
    np.testing.assert_array_equal(parrm.filter, g["filter"])
    out = parrm.filter_data()
    assert np.allclose(out, g["matlab_filtered"])


def test_example_data_registry():
    from pyparrm_b200.data import DATASETS

    for name in DATASETS:
        assert np.load(get_example_data_paths(name)).ndim == 2
    with pytest.raises(ValueError, match="`name` must be one of"):
        get_example_data_paths("nope")


def test_filter_sweep_host_logic(cpu_engine):
    """PARRM.filter_sweep: defaults, validation with create_filter's messages, per-set results
    equal to create_filter + filter_data one by one (stand-in engine; the GPU test covers the
    batched kernels)."""
    from pyparrm_b200 import PARRM
    from pyparrm_b200.synthetic import make_recording

    data = make_recording(2, 5000, 200, 13, seed=1)
    parrm = PARRM(data, 200, 13, verbose=False)
    with pytest.raises(ValueError, match="The period has not yet been estimated"):
        parrm.filter_sweep([{}])
    parrm._period = np.float64(200 / 13)
    sets = [dict(), dict(filter_half_width=500, filter_direction="past"),
            dict(omit_n_samples=3, period_half_width=0.5)]
    out, taps = parrm.filter_sweep(sets)
    assert out.shape == (3, 2, 5000) and parrm._filter is None   # object's own filter untouched
    for k, s in enumerate(sets):
        single = PARRM(data, 200, 13, verbose=False)
        single._period = parrm._period
        single.create_filter(**s)
        assert np.array_equal(taps[k], np.flatnonzero(single.filter < 0) - single._filter_half_width)
        assert np.array_equal(out[k], single.filter_data())
    for bad, exc, msg in (
        (dict(omit_n_samples=1.5), TypeError, "`omit_n_samples` must be an int."),
        (dict(omit_n_samples=-1), ValueError, "`omit_n_samples` must lie in the range"),
        (dict(period_half_width="a"), TypeError, "`period_half_width` must be an int or a float."),
        (dict(period_half_width=100.0), ValueError, "`period_half_width` must be lie in the range"),
        (dict(filter_half_width=2.0), TypeError, "`filter_half_width` must be an int."),
        (dict(filter_half_width=2, omit_n_samples=5), ValueError, "`filter_half_width` must lie in"),
        (dict(filter_direction=1), TypeError, "`filter_direction` must be a str."),
        (dict(filter_direction="up"), ValueError, "`filter_direction` must be one of"),
    ):
        with pytest.raises(exc, match=msg):
            parrm.filter_sweep([bad])


def test_compute_psd_wrapper_quirks(cpu_engine):
    """pyparrm_b200._utils._power.compute_psd: frequency axis, max_freq cut, and the
    reference's `psd[:-1] *= 2` (all ROWS but the last of a 2-D input, all BINS but the last of
    a 1-D input), against the oracle's restatement of _utils/_power.py:10-68."""
    from oracle import parrm_oracle as oracle
    from pyparrm_b200._utils._power import compute_psd

    rng = np.random.default_rng(44)
    for shape, fs, n, fmax in (((2, 100), 20, 10, None), ((3, 64), 100, 32, 30.0), ((50,), 20, 16, None)):
        x = rng.standard_normal(shape)
        freqs, psd = compute_psd(data=x, sampling_freq=fs, n_points=n, max_freq=fmax, n_jobs=2)
        f_want, p_want = oracle.periodogram(x, fs, n, fmax)
        assert np.array_equal(freqs, f_want) and psd.dtype == np.float32
        assert psd.shape == p_want.shape and np.allclose(psd, p_want, rtol=1e-6, atol=0)
