"""Lock-step, batched restatement of the 1-D Nelder-Mead search the reference runs.

The reference refines period estimates with ``scipy.optimize.fmin`` at its defaults
(``parrm.py:499-517``, ``:545-550``): SciPy's ``_minimize_neldermead`` with
``xatol = fatol = 1e-4``, ``maxiter = maxfun = 200 * N`` (N = 1), rho=1, chi=2, psi=0.5,
sigma=0.5 and an initial simplex ``{x0, 1.05 * x0}`` (SciPy 1.18 ``optimize/_optimize.py``).
Each objective evaluation there is one full harmonic fit; the chains are independent but each
is strictly sequential.

Here every chain is a small state machine.  One *round* asks the evaluator for all points any
chain could need in its current iteration -- reflection, expansion, outside and inside
contraction and the shrink point -- for all chains at once (one batched GPU launch), then
each chain consumes exactly the values SciPy's control flow would have looked at, with the
same comparisons, the same call counting (including the ``maxfun`` cut-off in mid-iteration)
and the same sort.  The simplex trajectory, ``x``, ``fval``, ``nit`` and ``nfev`` therefore
match ``fmin(..., full_output=True)`` exactly whenever the objective values match.
"""

from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

RHO, CHI, PSI, SIGMA = 1, 2, 0.5, 0.5
NONZDELT, ZDELT = 0.05, 0.00025


class _TooManyCalls(Exception):
    pass


class _Chain:
    def __init__(self, x0: float, xtol: float, ftol: float, maxiter: int, maxfun: int):
        x0 = np.float64(x0)
        self.sim = np.empty(2, dtype=np.float64)
        self.sim[0] = x0
        self.sim[1] = (1 + NONZDELT) * x0 if x0 != 0 else ZDELT
        self.fsim = np.full(2, np.inf, dtype=np.float64)
        self.xtol, self.ftol = xtol, ftol
        self.maxiter, self.maxfun = maxiter, maxfun
        self.fcalls = 0
        self.iterations = 1
        self.done = False

    # -- SciPy's call wrapper: refuse once maxfun evaluations have been spent
    def _spend(self) -> None:
        if self.fcalls >= self.maxfun:
            raise _TooManyCalls
        self.fcalls += 1

    def _sort(self) -> None:
        order = np.argsort(self.fsim)
        self.sim = np.take(self.sim, order, 0)
        self.fsim = np.take(self.fsim, order, 0)

    def start(self, values: np.ndarray) -> None:
        try:
            for k in range(2):
                self._spend()
                self.fsim[k] = values[k]
        except _TooManyCalls:
            pass
        self._sort()
        self._check_done()

    def _check_done(self) -> None:
        if not (self.fcalls < self.maxfun and self.iterations < self.maxiter):
            self.done = True
            return
        if (
            np.max(np.abs(self.sim[1:] - self.sim[0])) <= self.xtol
            and np.max(np.abs(self.fsim[0] - self.fsim[1:])) <= self.ftol
        ):
            self.done = True

    def proposals(self) -> np.ndarray:
        """Reflection, expansion, outside contraction, inside contraction, shrink point."""
        best, worst = self.sim[0], self.sim[1]
        xbar = best / 1  # centroid of all vertices but the worst (one vertex when N = 1)
        return np.array(
            [
                (1 + RHO) * xbar - RHO * worst,
                (1 + RHO * CHI) * xbar - RHO * CHI * worst,
                (1 + PSI * RHO) * xbar - PSI * RHO * worst,
                (1 - PSI) * xbar + PSI * worst,
                best + SIGMA * (worst - best),
            ],
            dtype=np.float64,
        )

    def advance(self, points: np.ndarray, values: np.ndarray) -> None:
        xr, xe, xc, xcc, xs = points
        fxr_v, fxe_v, fxc_v, fxcc_v, fxs_v = values
        sim, fsim = self.sim, self.fsim
        try:
            self._spend()
            fxr = fxr_v
            shrink = False
            if fxr < fsim[0]:
                self._spend()
                fxe = fxe_v
                if fxe < fxr:
                    sim[-1], fsim[-1] = xe, fxe
                else:
                    sim[-1], fsim[-1] = xr, fxr
            else:
                if fxr < fsim[-2]:  # second-worst is the best vertex when N = 1
                    sim[-1], fsim[-1] = xr, fxr
                else:
                    if fxr < fsim[-1]:
                        self._spend()
                        fxc = fxc_v
                        if fxc <= fxr:
                            sim[-1], fsim[-1] = xc, fxc
                        else:
                            shrink = True
                    else:
                        self._spend()
                        fxcc = fxcc_v
                        if fxcc < fsim[-1]:
                            sim[-1], fsim[-1] = xcc, fxcc
                        else:
                            shrink = True
                    if shrink:
                        sim[1] = xs
                        self._spend()
                        fsim[1] = fxs_v
            self.iterations += 1
        except _TooManyCalls:
            pass
        self._sort()
        self._check_done()

    @property
    def x(self) -> np.float64:
        return self.sim[0]

    @property
    def fval(self) -> np.float64:
        return np.min(self.fsim)


def fmin_batch(
    evaluate: Callable[[np.ndarray], np.ndarray],
    starts: Sequence[float],
    xtol: float = 1e-4,
    ftol: float = 1e-4,
    maxiter: int = 200,
    maxfun: int = 200,
):
    """Run one Nelder-Mead chain per entry of ``starts`` in lock step.

    ``evaluate`` maps a float64 vector of candidate periods to their objective values.
    Returns a list of ``(x, fval, iterations, fcalls)`` per chain, as
    ``scipy.optimize.fmin(..., full_output=True)[:4]`` would (with ``x`` a scalar).
    """
    chains = [_Chain(x0, xtol, ftol, maxiter, maxfun) for x0 in starts]
    if not chains:
        return []
    first = np.concatenate([c.sim for c in chains])
    values = np.asarray(evaluate(first), dtype=np.float64)
    for k, chain in enumerate(chains):
        chain.start(values[2 * k : 2 * k + 2])
    while True:
        active = [c for c in chains if not c.done]
        if not active:
            break
        points = [c.proposals() for c in active]
        values = np.asarray(evaluate(np.concatenate(points)), dtype=np.float64)
        for k, chain in enumerate(active):
            chain.advance(points[k], values[5 * k : 5 * k + 5])
    return [(c.x, c.fval, c.iterations, c.fcalls) for c in chains]
