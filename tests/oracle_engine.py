"""Stand-in for ``pyparrm_b200._engine.DeviceEngine`` backed by the CPU oracle.

TEST INFRASTRUCTURE.  It lets the host logic of ``pyparrm_b200.PARRM`` (validation, index
selection, grids, ranking, lock-step Nelder-Mead, candidate sharding) run under
``pytest -m "not gpu"`` in a container without a GPU.  The product never imports it.
"""

import os
from dataclasses import dataclass

import numpy as np

from oracle import parrm_oracle as oracle


@dataclass
class OracleTile:
    z: np.ndarray
    indices: np.ndarray
    n_indices: int
    n_chans: int


class OracleEngine:
    def __init__(self):
        self.evaluations = 0
        self.rounds = 0

    def prepare_tiles(self, data, index_sets, outlier_boundary, precision="fp64"):
        z = oracle.standardise(data, outlier_boundary)
        return [OracleTile(z, np.asarray(idx), len(idx), z.shape[0]) for idx in index_sets]

    def merge_channel_tiles(self, tile, all_gather, n_chans, per):
        import torch

        z = np.zeros((per, tile.z.shape[1]))
        z[: tile.n_chans] = tile.z
        z_all = all_gather(torch.from_numpy(z)).numpy().reshape(-1, z.shape[1])[:n_chans]
        return OracleTile(z_all, tile.indices, tile.n_indices, n_chans)

    def tile_from_standardised(self, z, indices):
        return OracleTile(np.asarray(z), np.asarray(indices), len(indices), z.shape[0])

    def standardise_full(self, data, outlier_boundary):
        return oracle.standardise(data, outlier_boundary)

    def evaluate(self, tile, periods, bandwidth, lambda_, n_chans_divisor):
        periods = np.asarray(periods, dtype=np.float64).ravel()
        self.evaluations += len(periods)
        self.rounds += 1
        return oracle.objective_many(
            periods, tile.z, tile.indices, bandwidth, lambda_, n_chans_divisor,
            n_jobs=min(8, os.cpu_count() or 1),
        ).astype(np.float64)

    def build_taps(self, period, phw, hw, omit, direction):
        return oracle.tap_offsets(period, phw, hw, omit, direction)

    def filter_host(self, data, taps, precision="fp64", strategy=None, out_dtype=None):
        out = oracle.apply_filter_direct(np.asarray(data, dtype=np.float64), taps)
        return out if out_dtype is None else out.astype(out_dtype)

    def filter_shard(self, chunk, taps, x0, t0, t1, n_total, precision="fp64"):
        return self.filter_host_window(chunk, taps, x0, t0, t1, n_total)

    def filter_host_window(self, chunk, taps, x0, t0, t1, n_total):
        """Time shard: zero-pad the chunk into place and keep the requested outputs."""
        chunk = np.asarray(chunk, dtype=np.float64)
        full = np.zeros((chunk.shape[0], n_total))
        full[:, x0:x0 + chunk.shape[1]] = chunk
        return oracle.apply_filter_direct(full, taps)[:, t0:t1]

    # ---- parameter sweeps / spectrum (host logic of PARRM.filter_sweep, compute_psd) -----------
    def default_half_widths(self, periods, period_half_widths, omits, limit):
        return np.array([oracle.default_half_width(p, w, int(o), 2 * int(limit) + 1)
                         for p, w, o in zip(periods, period_half_widths, omits)], dtype=np.int64)

    def filter_sweep(self, data, periods, period_half_widths, half_widths, omits, directions):
        data = np.asarray(data, dtype=np.float64)
        out = np.zeros((len(half_widths),) + data.shape)
        taps_all = []
        for k, (p, w, hw, om, d) in enumerate(zip(periods, period_half_widths, half_widths, omits,
                                                  directions)):
            taps = oracle.tap_offsets(p, w, int(hw), int(om), d)
            taps_all.append(taps)
            if len(taps):
                out[k] = oracle.apply_filter_direct(data, taps)
        return out, taps_all

    def periodogram(self, data, n_points, sampling_freq):
        x = np.asarray(data).astype(np.float32)
        coeffs = np.fft.fft(x, int(n_points))[..., 1:(int(n_points) // 2) + 1]
        return ((1.0 / (sampling_freq * n_points)) * np.abs(coeffs).astype(np.float32) ** 2).astype(
            np.float32)

