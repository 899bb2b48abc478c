"""CUDA period search (standardise, candidate evaluator, whole find_period) vs the reference.

Golden values come from the unmodified reference (tests/golden, oracle/make_golden.py).
Tolerance: relative error <= 1e-9 on every well-conditioned objective value and on the period
(BASELINE.json north_star).  Candidates whose Gram matrix the reference itself finds singular
or hopeless (objective inf or > 1e3; e.g. period exactly 4/3 samples on the bundled
recording) must come out non-finite or huge here too, never small.
"""

import numpy as np
import pytest

from oracle import parrm_oracle as oracle
from pyparrm_b200 import PARRM, get_example_data_paths
from pyparrm_b200.synthetic import make_recording

pytestmark = pytest.mark.gpu
RTOL = 1e-9
HUGE = 1e3


# candidates that needed the "numerically singular" escape hatch (cond > 1e6, 1e-4), so that a
# regression pushing well-conditioned candidates into it shows up in the test output
LOOSE_BRANCH = {"count": 0, "compared": 0}


def compare_objective(got, want, label="", periods=None, indices=None, bandwidth=None):
    """Relative error <= 1e-9 on every candidate the reference can itself resolve.

    A candidate whose Gram matrix is numerically singular (condition number > 1e6; e.g. a period
    of exactly 31/2 samples gives 31 distinct phases for 41 unknowns) has no value that is
    defined to 1e-9 in the reference either -- LAPACK happens not to meet an exact zero pivot
    and returns rounding-dependent coefficients.  Those only have to agree loosely (1e-4), or
    be huge / non-finite on both sides.
    """
    want = np.asarray(want, dtype=np.float64)
    good = np.isfinite(want) & (want < HUGE)
    rel = np.zeros_like(want)
    rel[good] = np.abs(got[good] - want[good]) / np.abs(want[good])
    worst = 0.0
    for k in np.flatnonzero(good):
        if rel[k] <= RTOL:
            worst = max(worst, float(rel[k]))
            continue
        LOOSE_BRANCH["count"] += 1
        assert periods is not None, f"{label}: candidate {k} rel err {rel[k]:.3e}"
        design = oracle.harmonic_design(np.asarray(indices), periods[k], int(bandwidth))
        cond = np.linalg.cond(design.T @ design)
        assert cond > 1e6 and rel[k] <= 1e-4, (
            f"{label}: period {periods[k]!r} rel err {rel[k]:.3e}, Gram condition {cond:.2e}")
    bad = ~good
    assert np.all(~np.isfinite(got[bad]) | (got[bad] > HUGE)), f"{label}: degenerate candidates"
    LOOSE_BRANCH["compared"] += int(good.sum())
    return worst


def test_standardise(gpu_engine):
    import torch

    for dtype, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        x = make_recording(5, 70_001, 2000, 130, seed=21).astype(dtype)
        x[3] *= 1e-3
        z = oracle.standardise(x, 3.0)
        idx_sets = [np.arange(100, 5101), np.unique(np.random.default_rng(0).integers(0, 69_000, 20_000)) + 500]
        tiles = gpu_engine.prepare_tiles(x, idx_sets, 3.0)
        for tile, idx in zip(tiles, idx_sets):
            y = tile.y.cpu().numpy()
            assert y.shape == (idx.shape[0], 5)
            assert np.abs(y - z[:, idx].T).max() <= tol * 3.0
            assert np.allclose(tile.sumsq.cpu().numpy(), (z[:, idx].astype(np.float64) ** 2).sum(1), rtol=max(tol, 1e-12) * 10)
        full = gpu_engine.standardise_full(x, 3.0)
        assert full.dtype == dtype and full.shape == z.shape
        assert np.abs(full - z).max() <= tol * 3.0


def test_objective_golden_cases(golden, gpu_engine):
    g = golden("objective")
    n, fs, fa = (int(v) for v in g["recording"])
    worst = 0.0
    for case in range(int(g["n_cases"])):
        n_chans, bw, lam, seed = g[f"case{case}_params"]
        data = make_recording(int(n_chans), n, fs, fa, seed=int(seed))
        idx = g[f"case{case}_indices"]
        (tile,) = gpu_engine.prepare_tiles(data, [idx], 3.0)
        got = gpu_engine.evaluate(tile, g[f"case{case}_periods"], int(bw), float(lam), int(n_chans))
        worst = max(worst, compare_objective(got, g[f"case{case}_values"], f"case {case}",
                                             g[f"case{case}_periods"], idx, bw))
    print(f"objective golden: worst relative error {worst:.3e}")


@pytest.mark.parametrize("name", ["example_dbs", "synthetic_2x30000", "ecog_lfp"])
def test_every_recorded_evaluation(golden, gpu_engine, name):
    """All ~1 750 objective evaluations the reference made during find_period, stage by stage."""
    g = golden(name)
    if name == "example_dbs":
        data = np.load(get_example_data_paths("example_data"))
    elif name == "ecog_lfp":
        data = np.load(get_example_data_paths("ecog_lfp_data"))
    else:
        n_chans, n, fs, fa, seed = (int(v) for v in g["recording"])
        data = make_recording(n_chans, n, fs, fa, seed=seed)
    idx_sets = [g[f"run{r}_indices"] for r in range(int(g["n_runs"]))]
    tiles = gpu_engine.prepare_tiles(data, idx_sets, 3.0)
    calls = g["calls"]
    by_len = {len(i): t for i, t in zip(idx_sets, tiles)}
    worst = 0.0
    stages = sorted({(int(b), float(l), int(n)) for _, b, l, n, _ in calls})
    for bw, lam, n_idx in stages:
        rows = calls[(calls[:, 1] == bw) & (calls[:, 2] == lam) & (calls[:, 3] == n_idx)]
        got = gpu_engine.evaluate(by_len[n_idx], rows[:, 0], bw, lam, data.shape[0])
        idx = idx_sets[[len(i) for i in idx_sets].index(n_idx)]
        worst = max(worst, compare_objective(got, rows[:, 4], f"{name} bw={bw} lambda={lam}",
                                             rows[:, 0], idx, bw))
    print(f"{name}: {len(calls)} evaluations, worst relative error {worst:.3e}; so far "
          f"{LOOSE_BRANCH['count']} of {LOOSE_BRANCH['compared']} compared candidates needed the "
          "cond > 1e6 branch (1e-4)")
    # the escape hatch is for a handful of numerically singular candidates, not a tolerance
    assert LOOSE_BRANCH["count"] <= 0.01 * LOOSE_BRANCH["compared"] + 5


def test_batch_split_and_single_candidate_agree(gpu_engine):
    """Few candidates (sample-split CTAs) and many candidates (one CTA each) give the same value."""
    data = make_recording(3, 60_000, 2000, 130, seed=12)
    idx = np.arange(10_000, 35_001)
    (tile,) = gpu_engine.prepare_tiles(data, [idx], 3.0)
    periods = 2000 / 130 * (1 + np.linspace(-1e-3, 1e-3, 700))
    many = gpu_engine.evaluate(tile, periods, 20, 1.0, 3)
    for k in (0, 123, 699):
        one = gpu_engine.evaluate(tile, periods[k : k + 1], 20, 1.0, 3)
        assert abs(one[0] - many[k]) <= 1e-12 * abs(many[k])
    val, pos = gpu_engine.argmin(gpu_engine.evaluate_device(tile, periods, 20, 1.0, 3))
    assert pos == int(np.argmin(many)) and val == many.min()


def test_wide_and_narrow_channel_counts(gpu_engine):
    """1, 64 (one full channel tile), 65 and 130 channels against the oracle."""
    base = make_recording(130, 12_000, 2000, 130, seed=13)
    idx = np.arange(1000, 6001)
    periods = 2000 / 130 * (1 + np.array([-2e-3, 0.0, 3e-6, 1e-3]))
    for n_chans in (1, 64, 65, 130):
        data = np.ascontiguousarray(base[:n_chans])
        (tile,) = gpu_engine.prepare_tiles(data, [idx], 3.0)
        got = gpu_engine.evaluate(tile, periods, 10, 1.0, n_chans)
        z = oracle.standardise(data, 3.0)
        want = oracle.objective_many(periods, z, idx, 10, 1.0, n_chans, n_jobs=8)
        compare_objective(got, want, f"{n_chans} channels", periods, idx, 10)


@pytest.mark.parametrize("bandwidth", [1, 4, 7, 12, 16, 23])
def test_every_row_block_count_and_tile_layout(gpu_engine, bandwidth):
    """Bandwidths that need 1..6 eight-row blocks of W' (the tensor kernel is compiled per block
    count), on an odd channel count (re-tiled copy of Y) and a small even one (the caller's
    array read directly), with a ragged last tile of samples."""
    base = make_recording(6, 9_000, 2000, 130, seed=17)
    idx = np.arange(700, 700 + 128 * 9 + 37)
    periods = 2000 / 130 * (1 + np.array([-1.5e-3, 2e-6, 7e-4]))
    for n_chans in (3, 6):
        data = np.ascontiguousarray(base[:n_chans])
        (tile,) = gpu_engine.prepare_tiles(data, [idx], 3.0)
        got = gpu_engine.evaluate(tile, periods, bandwidth, 1.0, n_chans)
        z = oracle.standardise(data, 3.0)
        want = oracle.objective_many(periods, z, idx, bandwidth, 1.0, n_chans, n_jobs=4)
        compare_objective(got, want, f"bw {bandwidth}, {n_chans} channels", periods, idx, bandwidth)


@pytest.mark.parametrize("n_fit", [40, 128, 129, 512, 640, 1153])
def test_short_and_tile_aligned_sample_counts(gpu_engine, n_fit):
    """Fewer samples than one 128-sample tile, exact tile / batch multiples and one past them,
    through the tensor kernel (4 channels) and the narrow one (1 channel)."""
    base = make_recording(4, 6_000, 2000, 130, seed=19)
    idx = np.arange(300, 300 + n_fit)
    periods = 2000 / 130 * (1 + np.array([-1e-3, 4e-6, 2e-3, 5e-3]))
    for n_chans, bandwidth in ((4, 2), (4, 5), (1, 5)):
        data = np.ascontiguousarray(base[:n_chans])
        (tile,) = gpu_engine.prepare_tiles(data, [idx], 3.0)
        got = gpu_engine.evaluate(tile, periods, bandwidth, 1.0, n_chans)
        z = oracle.standardise(data, 3.0)
        want = oracle.objective_many(periods, z, idx, bandwidth, 1.0, n_chans, n_jobs=4)
        compare_objective(got, want, f"N {n_fit}, {n_chans} ch, bw {bandwidth}", periods, idx,
                          bandwidth)


@pytest.mark.parametrize("name", ["example_dbs", "synthetic_2x30000", "ecog_lfp"])
def test_find_period_matches_reference(golden, gpu_engine, name):
    g = golden(name)
    if name == "example_dbs":
        parrm = PARRM(np.load(get_example_data_paths("example_data")), 200, 150, verbose=False)
        parrm.find_period()
    elif name == "ecog_lfp":
        parrm = PARRM(np.load(get_example_data_paths("ecog_lfp_data")), 1000, 130, verbose=False)
        parrm.find_period(random_seed=0)
    else:
        n_chans, n, fs, fa, seed = (int(v) for v in g["recording"])
        parrm = PARRM(make_recording(n_chans, n, fs, fa, seed=seed), fs, fa, verbose=False)
        parrm.find_period(random_seed=0)
    want = float(g["period"])
    print(f"{name}: period {parrm.period!r} reference {want!r}")
    assert isinstance(parrm.period, np.float64)
    assert abs(parrm.period - want) <= RTOL * want


def test_whole_workflow_known_answer(golden, gpu_engine):
    """examples/plot_use_parrm.py end to end: period, filter, np.allclose(matlab_filtered)."""
    g = golden("example_dbs")
    data = np.load(get_example_data_paths("example_data"))
    parrm = PARRM(data=data, sampling_freq=200, artefact_freq=150, verbose=False)
    parrm.find_period()
    parrm.create_filter(filter_half_width=2000, omit_n_samples=20, filter_direction="both",
                        period_half_width=0.01)
    out = parrm.filter_data()
    assert np.array_equal(np.flatnonzero(parrm.filter < 0) - 2000, g["taps"])
    assert np.allclose(out, g["matlab_filtered"])
    assert np.abs(out - g["filtered"]).max() <= RTOL * np.abs(data).max()


@pytest.mark.parametrize("n_chans", [1, 2])
@pytest.mark.parametrize("n_samples", [100, 300, 25000])
@pytest.mark.parametrize("search_portion", [None, 0.5])
def test_reference_suite_matrix(gpu_engine, n_chans, n_samples, search_portion):
    """The reference's own smoke matrix (tests/test_parrm.py:17-68) incl. degenerate period 2.0."""
    rng = np.random.default_rng(44)
    data = rng.standard_normal((n_chans, n_samples))
    parrm = PARRM(data=data, sampling_freq=20, artefact_freq=10, verbose=False)
    search = None if search_portion is None else np.arange(0, n_samples * search_portion)
    parrm.find_period(search_samples=search, assumed_periods=20 / 10, random_seed=44, n_jobs=2)
    for direction in ["future", "past", "both"]:
        parrm.create_filter(filter_direction=direction)
    filtered = parrm.filter_data()
    assert filtered.shape == data.shape and isinstance(filtered, np.ndarray)
    other = rng.standard_normal((1, 50))
    assert parrm.filter_data(other).shape == other.shape
    assert repr(parrm) == (
        f"PARRM object | Data: ({n_chans} channels x {n_samples} times) | Period: {parrm.period :.4f}")
    assert np.isfinite(parrm.period)


def test_large_sample_indices(gpu_engine):
    """Phase angles of a long recording (indices ~1e6, angles ~5e5 rad) go through the
    evaluator's own argument reduction instead of sincos()'s slow path: check it against the
    oracle there, for both the contiguous and the random-index branch."""
    n_chans, n = 2, 1_200_000
    data = make_recording(n_chans, n, 2000, 130, seed=21)
    z = oracle.standardise(data, 3.0)
    rng = np.random.default_rng(3)
    per = 2000 / 130 * (1 + 3e-6)
    periods = per * (1 + np.array([-2e-3, -1e-5, 0.0, 3e-6, 4e-4]))
    for idx in (np.arange(1_100_000, 1_105_001),
                np.unique(rng.integers(0, n - 60_002, 6000)) + 30_000):
        (tile,) = gpu_engine.prepare_tiles(data, [idx], 3.0)
        got = gpu_engine.evaluate(tile, periods, 20, 1.0, n_chans)
        want = oracle.objective_many(periods, z, idx, 20, 1.0, n_chans, n_jobs=4)
        compare_objective(got, want, "large indices", periods, idx, 20)


def test_device_nelder_mead_retraces_host_state_machines(gpu_engine):
    """csrc/neldermead.cu vs pyparrm_b200/_neldermead.py (itself pinned to scipy.optimize.fmin
    by tests/test_neldermead.py): x, fval, nit and nfev of every chain bit-equal, when both
    see the same objective values (same evaluator, same batch shape of 5 points per chain)."""
    from pyparrm_b200._neldermead import fmin_batch

    data = make_recording(3, 40_000, 2000, 130, seed=12)
    z = oracle.standardise(data, 3.0)
    for idx, bw, lam, starts in (
        (np.arange(5_000, 10_001), 5, 1.0, [15.38, 15.3846, 15.40, 15.2, 15.39]),
        (np.arange(2_000, 27_001), 20, 0.0, [2000 / 130]),
        (np.arange(100, 1_100), 10, 1.0, [7.7, 15.5]),
    ):
        tile = gpu_engine.tile_from_standardised(z, idx)
        n = len(starts)

        def evaluate(p, tile=tile, n=n, bw=bw, lam=lam):
            padded = np.resize(np.asarray(p, dtype=np.float64), 5 * n)  # same launch shape
            return gpu_engine.evaluate(tile, padded, bw, lam, 3)[: len(p)]

        want = fmin_batch(evaluate, starts)
        launches0 = gpu_engine.launches
        got = gpu_engine.nm_minimise(tile, starts, bw, lam, 3)
        rounds = max(w[2] for w in want)
        # init + first round eagerly (<= 7 kernels), then one launch per replayed graph
        assert gpu_engine.launches - launches0 <= 8 + rounds // gpu_engine.ROUNDS_PER_GRAPH + 1
        for g, w in zip(got, want):
            assert g[0] == w[0] and (g[1] == w[1] or (np.isnan(g[1]) and np.isnan(w[1])))
            assert g[2:] == (w[2], w[3]), (g, w)


@pytest.mark.parametrize("name", ["example_dbs", "synthetic_2x30000", "ecog_lfp"])
def test_fp32_mode_period_within_1e_4(golden, gpu_engine, name):
    """BASELINE north_star: the fp32 mode reproduces the reference period to <= 1e-4 relative.
    Search tiles are stored in float32 (parrm_eval_periods_typed); the fit runs in float64."""
    import torch

    g = golden(name)
    if name == "example_dbs":
        parrm = PARRM(np.load(get_example_data_paths("example_data")), 200, 150, verbose=False,
                      precision="fp32")
        parrm.find_period()
    elif name == "ecog_lfp":
        parrm = PARRM(np.load(get_example_data_paths("ecog_lfp_data")), 1000, 130, verbose=False,
                      precision="fp32")
        parrm.find_period(random_seed=0)
    else:
        n_chans, n, fs, fa, seed = (int(v) for v in g["recording"])
        parrm = PARRM(make_recording(n_chans, n, fs, fa, seed=seed), fs, fa, verbose=False,
                      precision="fp32")
        parrm.find_period(random_seed=0)
    want = float(g["period"])
    rel = abs(parrm.period - want) / want
    print(f"{name}: fp32-mode period {parrm.period!r} reference {want!r} rel {rel:.2e}")
    assert rel <= 1e-4
    # the evaluator itself: float32 tile vs float64 tile on one grid
    data = make_recording(5, 30_000, 2000, 130, seed=3)
    idx = np.arange(2_000, 12_001)
    t64, = gpu_engine.prepare_tiles(data, [idx], 3.0)
    t32, = gpu_engine.prepare_tiles(data, [idx], 3.0, precision="fp32")
    assert t32.y.dtype == torch.float32 and t64.y.dtype == torch.float64
    periods = (2000 / 130) * (1 + np.linspace(-1e-3, 1e-3, 41))
    e64 = gpu_engine.evaluate(t64, periods, 10, 1.0, 5)
    e32 = gpu_engine.evaluate(t32, periods, 10, 1.0, 5)
    assert np.abs(e32 - e64).max() <= 1e-4 * np.abs(e64).max()
    assert int(np.argmin(e32)) == int(np.argmin(e64))


def test_cfg2_full_size_against_reference(golden, gpu_engine):
    """BASELINE cfg2 at its stated size (64 channels x 1.2 M samples): the objective on the
    run-3 shape for 16 grid candidates, and the filtered output, against values recorded from
    the unmodified reference (tests/golden/cfg2_objective.npz, oracle/make_golden.py cfg2)."""
    g = golden("cfg2_objective")
    n_chans, n, fs, fa, seed = (int(v) for v in g["recording"])
    data = make_recording(n_chans, n, fs, fa, seed=seed)
    tile, = gpu_engine.prepare_tiles(data, [g["indices"]], 3.0)
    got = gpu_engine.evaluate(tile, g["periods"], 20, 1.0, n_chans)
    worst = compare_objective(got, g["values"], "cfg2 run-3 shape", g["periods"], g["indices"], 20)
    print(f"cfg2 full size: worst objective rel err {worst:.2e}")
    parrm = PARRM(data[:2], fs, fa, verbose=False)
    parrm._period = np.float64(fs / fa * (1 + 3e-6))
    parrm.create_filter(filter_half_width=2000, filter_direction="both")
    y = parrm.filter_data()
    scale = np.abs(data[:2]).max()
    assert np.abs(y[1, :4000] - g["filtered_ch1_head"]).max() <= RTOL * scale
    assert np.abs(y[1, 600000:602000] - g["filtered_ch1_mid"]).max() <= RTOL * scale
