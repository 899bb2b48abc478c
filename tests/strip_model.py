"""NumPy model of the comb-box strip kernel (pyparrm_b200/csrc/filter.cu).

Test infrastructure only.  It restates, step by step and with the same ring/mirror index
arithmetic, what ``filter_comb_strip_kernel`` does for one channel, so that the design (ring
slots, mirror chunk, sliding comb boxes, piece boundaries, edge counts) is checked against
the oracle on the CPU; the CUDA kernel itself is checked against the oracle on the GPU.
"""

from __future__ import annotations

import numpy as np


def _ceil_div(a, b):
    return -(-a // b)


class StripModel:
    def __init__(self, taps, desc, tile, prefetch=2, reinit_every=0, threads=256, pipe=False,
                 stages=2):
        self.taps = np.asarray(taps, dtype=np.int64)
        self.n_taps = len(self.taps)
        self.w_lo = min(int(self.taps[0]), 0)
        self.w_hi = max(int(self.taps[-1]), 0)
        self.tile = tile
        self.P = prefetch
        self.pipe = pipe
        self.stages = stages           # pipelined kernel: the slide runs stages - 1 chunks ahead
        self.extra = stages if pipe else 1  # ring chunks beyond what one step reads
        self.reinit_every = reinit_every
        self.d = desc["stride"]
        self.nk = len(desc["windows"])
        self.m = list(desc["windows"])
        self.boxes = [np.asarray(b, dtype=np.int64) for b in desc["boxes"]]
        self.plus = np.asarray(desc["plus"], dtype=np.int64)
        self.minus = np.asarray(desc["minus"], dtype=np.int64)
        self.centre = desc["centre"]
        self.h_back = _ceil_div(self.w_hi + self.d, tile)
        self.h_fwd = _ceil_div(-self.w_lo, tile)
        self.nq_x = self.h_back + self.h_fwd + self.extra + self.P
        self.RX = self.nq_x * tile
        self.a_lo = [int(b.min()) for b in self.boxes]
        self.nq_d = [
            _ceil_div(max(int(b.max()) - int(b.min()), self.d), tile) + self.extra for b in self.boxes
        ]
        self.RD = [q * tile for q in self.nq_d]
        self.cx1 = [(-a) % self.RX for a in self.a_lo]
        self.cx2 = [(-(a + m * self.d)) % self.RX for a, m in zip(self.a_lo, self.m)]
        self.cprev = [(-self.d) % rd for rd in self.RD]
        self.off_box = [(-(b - a)) % rd for b, a, rd in zip(self.boxes, self.a_lo, self.RD)]
        self.off_plus = (-self.plus) % self.RX
        self.off_minus = (-self.minus) % self.RX
        # D-pass work split (launch_strip): one item per chain unless there are few chains
        self.seg_len, self.n_seg = [], []
        for k in range(self.nk):
            chains, per_chain = min(self.d, tile), _ceil_div(tile, self.d)
            seg = per_chain
            if chains * self.nk < threads / 2:
                seg = min(max(9, self.m[k]) | 1, per_chain)
            self.seg_len.append(seg)
            self.n_seg.append(_ceil_div(per_chain, seg))

    # ------------------------------------------------------------------
    def run(self, x, x_t0, t0, n_out, n_total, gamma=0, n_pieces=1):
        """x holds global times [x_t0, x_t0 + len(x)); returns outputs for [t0, t0 + n_out)."""
        tile = self.tile
        out = np.full(n_out, np.nan)
        n_x = len(x)
        self.lo_valid = max(0, x_t0)
        self.hi_valid = min(n_total, x_t0 + n_x)
        self.x, self.x_t0, self.n_total = x, x_t0, n_total
        steps_per_chan = _ceil_div(n_out + tile - 1, tile)
        j_first = (t0 - gamma) // tile
        j_last = (t0 + n_out - 1 - gamma) // tile
        for piece in range(n_pieces):
            s0 = steps_per_chan * piece // n_pieces
            s1 = steps_per_chan * (piece + 1) // n_pieces
            js0, js_end = j_first + s0, min(j_first + s1, j_last + 1)
            if js0 >= js_end:
                continue
            self._piece(js0, js_end, gamma, t0, n_out, out)
        return out

    def _load_chunk(self, j, js0, gamma):
        tile = self.tile
        slot = (j - js0 + self.h_back) % self.nq_x
        g = gamma + j * tile + np.arange(tile)
        ok = (g >= self.lo_valid) & (g < self.hi_valid)
        v = np.zeros(tile)
        v[ok] = self.x[g[ok] - self.x_t0]
        self.sX[slot * tile:(slot + 1) * tile] = v
        if slot == 0:
            self.sX[self.RX:self.RX + tile] = v

    def _init_boxes(self, n):
        tile = self.tile
        sxn = ((n + self.h_back) % self.nq_x) * tile
        for k in range(self.nk):
            back = self.nq_d[k] - self.extra
            n_back = back * tile
            e = np.arange(n_back)
            rel_i = -self.a_lo[k] - n_back + e
            total = np.zeros(n_back)
            for q in range(self.m[k]):
                rel = rel_i - q * self.d
                ok = rel >= -self.h_back * tile
                pos = (sxn + rel) % self.RX  # wrap_both
                total[ok] += self.sX[pos[ok]]
            chunk, within = e // tile, e % tile
            slot = (n + chunk) % self.nq_d[k]
            self.sD[k][slot * tile + within] = total
            first = slot == 0
            self.sD[k][self.RD[k] + within[first]] = total[first]

    def _piece(self, js0, js_end, gamma, t0, n_out, out):
        tile = self.tile
        self.sX = np.full(self.RX + tile, np.nan)
        self.sD = [np.full(rd + tile, np.nan) for rd in self.RD]
        n_steps = js_end - js0
        j_need_max = js_end - 1 + self.h_fwd
        for j in range(js0 - self.h_back, js0 + self.h_fwd + 1):
            self._load_chunk(j, js0, gamma)
        pending = set()

        def issue(j):
            # an asynchronous chunk lands at the latest legal moment (its wait); until then
            # its ring slot holds NaN, so both early reads and premature reuse show up
            if j <= j_need_max:
                pending.add(j)
                slot = (j - js0 + self.h_back) % self.nq_x
                self.sX[slot * tile:(slot + 1) * tile] = np.nan
                if slot == 0:
                    self.sX[self.RX:] = np.nan

        def land(j):
            if j in pending:
                pending.discard(j)
                self._load_chunk(j, js0, gamma)

        for j in range(js0 + self.h_fwd + 1, js0 + self.h_fwd + self.P):
            issue(j)
        self._init_boxes(0)
        if not self.pipe:
            for n in range(n_steps):
                js = js0 + n
                issue(js + self.h_fwd + self.P)
                if self.reinit_every and n > 0 and n % self.reinit_every == 0:
                    self._init_boxes(n)
                self._slide(n)
                self._gather(n, gamma + js * tile, t0, n_out, out)
                land(js + self.h_fwd + 1)
        else:
            # the slide warps run as far ahead as the hand-off barriers allow: slide(n + stages - 1)
            # (and its TMA issue) completes before gather(n) starts
            def slide_step(n):
                issue(js0 + n + self.h_fwd + self.P)
                self._slide(n)
                land(js0 + n + 1 + self.h_fwd)

            ahead = self.stages - 1
            for m in range(min(ahead, n_steps)):
                slide_step(m)
            for n in range(n_steps):
                if n + ahead < n_steps:
                    slide_step(n + ahead)
                self._gather(n, gamma + (js0 + n) * tile, t0, n_out, out)

    def _slide(self, n):
        tile, d = self.tile, self.d
        sxn = ((n + self.h_back) % self.nq_x) * tile
        for k in range(self.nk):
            back = self.nq_d[k] - self.extra
            slot = (n + back) % self.nq_d[k]
            base1 = (sxn + self.cx1[k]) % self.RX
            base2 = (sxn + self.cx2[k]) % self.RX
            base_prev = (slot * tile + self.cprev[k]) % self.RD[k]
            chains = min(d, tile)
            L = self.seg_len[k]
            u = np.arange(chains * self.n_seg[k])
            s_idx, c = u // chains, u % chains
            e = c + s_idx * L * d
            total = np.zeros(len(u))
            first = s_idx == 0
            total[first] = self.sD[k][base_prev + c[first]]
            for q in range(1, self.m[k] + 1):  # direct D[i - d] for the later segments
                rest = ~first & (e < tile)
                total[rest] += self.sX[(base1 + e[rest] - q * d) % self.RX]
            new = {}
            for r in range(L):
                live = e < tile
                el = e[live]
                total[live] += self.sX[base1 + el] - self.sX[base2 + el]
                new.update(zip(el.tolist(), total[live].tolist()))
                e = e + d
            assert len(new) == tile  # every element of the chunk written exactly once
            el = np.fromiter(new.keys(), dtype=np.int64)
            vals = np.fromiter(new.values(), dtype=np.float64)
            self.sD[k][slot * tile + el] = vals
            if slot == 0:
                self.sD[k][self.RD[k] + el] = vals

    def _gather(self, n, cur, t0, n_out, out):
        tile = self.tile
        sxn = ((n + self.h_back) % self.nq_x) * tile
        i = np.arange(tile)
        acc = np.zeros(tile)
        for k in range(self.nk):
            back = self.nq_d[k] - self.extra
            sdn = ((n + back) % self.nq_d[k]) * tile
            for off in self.off_box[k]:
                base = (sdn + off) % self.RD[k]
                acc += self.sD[k][base + i]
        for off in self.off_plus:
            acc += self.sX[(sxn + off) % self.RX + i]
        for off in self.off_minus:
            acc -= self.sX[(sxn + off) % self.RX + i]
        xc = self.sX[sxn + i]
        total = acc + self.centre * xc
        g = cur + i
        interior = (cur - self.w_hi >= 0) and (cur + tile - self.w_lo <= self.n_total)
        if interior:
            y = xc - total * (1.0 / self.n_taps)
        else:
            n_in = (np.searchsorted(self.taps, g, side="right")
                    - np.searchsorted(self.taps, g - self.n_total, side="right"))
            with np.errstate(divide="ignore", invalid="ignore"):
                y = np.where(n_in > 0, xc - total / n_in, 0.0)
        keep = (g >= t0) & (g < t0 + n_out)
        out[g[keep] - t0] = y[keep]
