"""NumPy restatement of the index arithmetic of the run-time specialised filter kernel
(pyparrm_b200/csrc/filter_comb_e.cuh), for CPU tests of the design: chunk grid and alignment
shift, strips ("pieces") with one priming block, register rings of an unrolled block that is
a whole number of chunks, sliding sums re-added once per block, recording edges through the
plan's cumulative tap-count table, partial output ranges, and the non-finite rule.

One Python loop iteration = one comb step of all chains (vectorised over the chains)."""

from __future__ import annotations

import numpy as np


class PatternModel:
    def __init__(self, taps, desc, steps_per_chunk):
        assert desc["kind"] == 1
        self.taps = np.asarray(taps, dtype=np.int64)
        self.d = int(desc["stride"])
        wins = list(desc["windows"])
        boxes = [np.asarray(b, dtype=np.int64) for b in desc["boxes"]]
        if len(wins) == 2 and wins[1] > wins[0]:  # longer box first, as filter_jit.cu orders them
            wins, boxes = wins[::-1], boxes[::-1]
        self.m, self.off = wins, boxes
        self.plus = np.asarray(desc["plus"], dtype=np.int64)
        self.minus = np.asarray(desc["minus"], dtype=np.int64)
        self.centre = int(desc["centre"])
        self.n_taps = len(self.taps)
        self.w_lo, self.w_hi = min(int(self.taps[0]), 0), max(int(self.taps[-1]), 0)
        self.u = steps_per_chunk
        self.ch = self.u * self.d
        self.gpb = -(-self.m[0] // self.u)       # groups per unrolled block = priming groups
        self.b = self.gpb * self.u               # steps per block (>= M0)
        # plan tables
        self.count = np.searchsorted(self.taps, np.arange(self.w_lo - 1, self.w_hi + 1), "right")
        self.recip = np.concatenate([[0.0], 1.0 / np.arange(1, self.n_taps + 1)])

    def _edge(self, t, n_total, xc, tot):
        v1 = np.clip(t, self.w_lo - 1, self.w_hi)
        v2 = np.clip(t - n_total, self.w_lo - 1, self.w_hi)
        n_in = self.count[v1 - (self.w_lo - 1)] - self.count[v2 - (self.w_lo - 1)]
        y = xc - tot * self.recip[n_in]
        return np.where(n_in == 0, 0.0, y)

    def run(self, x, x_t0, t0, n_out, n_total, gamma=0, n_pieces=1):
        """x holds global times [x_t0, x_t0 + len(x)); returns outputs [t0, t0 + n_out)."""
        d, ch, u = self.d, self.ch, self.u
        lo_valid, hi_valid = max(0, x_t0), min(n_total, x_t0 + len(x))

        def sample(g):  # zero outside the available samples (the producer's zero fill)
            g = np.asarray(g)
            ok = (g >= lo_valid) & (g < hi_valid)
            return np.where(ok, x[np.clip(g - x_t0, 0, len(x) - 1)], 0.0)

        out = np.full(n_out, np.nan)
        j_first = (t0 - gamma) // ch
        j_last = (t0 + n_out - 1 - gamma) // ch
        gpc = j_last - j_first + 1
        c = np.arange(d)
        with np.errstate(invalid="ignore", over="ignore"):
            for piece in range(n_pieces):
                s0, s1 = gpc * piece // n_pieces, gpc * (piece + 1) // n_pieces
                if s0 >= s1:
                    continue
                n_groups = self.gpb + (s1 - s0)
                T0 = gamma + (j_first + s0 - self.gpb) * ch
                rings = [np.zeros((self.b, d)) for _ in self.m]
                g = 0
                while g < n_groups:
                    sums = [sum(r[(self.b - 1 - i) % self.b] for i in range(m))
                            for r, m in zip(rings, self.m)]
                    Tb = T0 + g * ch
                    for s in range(self.b):
                        gi = s // u
                        if g + gi >= n_groups:
                            break
                        t = Tb + s * d + c
                        single = np.zeros(d)
                        for k, (r, m, off) in enumerate(zip(rings, self.m, self.off)):
                            sums[k] = sums[k] - r[(s + self.b - m) % self.b]
                            e = sample(t - off[0])
                            if len(off) > 1:
                                eb = sample(t - off[1])
                                for bi in range(2, len(off)):
                                    if bi & 1:
                                        eb = eb + sample(t - off[bi])
                                    else:
                                        e = e + sample(t - off[bi])
                                e = e + eb
                            r[s] = e
                            sums[k] = sums[k] + e
                        for w in self.plus:
                            single = single + sample(t - w)
                        for w in self.minus:
                            single = single - sample(t - w)
                        xc = sample(t)
                        single = single + self.centre * xc
                        if g + gi < self.gpb:
                            continue  # priming group: rings only
                        Tn = Tb + gi * ch
                        interior = (Tn - self.w_hi >= 0) and (Tn + ch - self.w_lo <= n_total)

                        def value(tot):
                            if interior:
                                return xc - tot / self.n_taps
                            return self._edge(t, n_total, xc, tot)

                        y = value(sum(sums) + single)
                        bad = ~np.isfinite(y)
                        if bad.any():
                            # sums re-added from the rings for the chains that need it; the
                            # flagged outputs are re-evaluated tap by tap (direct_value)
                            for k, (r, m) in enumerate(zip(rings, self.m)):
                                fresh = sum(r[(s + self.b - i) % self.b] for i in range(m))
                                sums[k] = np.where(bad, fresh, sums[k])
                            src = t[bad][:, None] - self.taps[None, :]
                            inside = (src >= 0) & (src < n_total)
                            acc = np.where(inside, sample(src), 0.0).sum(axis=1)
                            n_in = inside.sum(axis=1)
                            y2 = np.where(n_in > 0, xc[bad] - acc * self.recip[n_in], 0.0)
                            y[bad] = np.where(np.isfinite(y2), y2, 0.0)
                        keep = (t >= t0) & (t < t0 + n_out)
                        out[t[keep] - t0] = y[keep]
                    g += self.gpb
        return out
