"""Development check of the run-time specialised filter kernel: parity vs the oracle on small
recordings (all kernels), then timing at the cfg2 shape.  python scripts/check_comb_e.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import parrm_oracle as oracle  # noqa: E402
from pyparrm_b200 import _native  # noqa: E402
from pyparrm_b200._engine import get_engine  # noqa: E402

eng = get_engine()
K = _native
CASES = {
    "cfg2": (2000 / 130 * (1 + 3e-6), None, 2000, 0, "both"),
    "cfg3": (1000 / 145 * (1 + 3e-6), None, 2469, 0, "both"),
    "cfg1": (1.3311148014466094, 0.01, 2000, 20, "both"),
    "cfg2past": (2000 / 130 * (1 + 3e-6), None, 2000, 0, "past"),
    "cfg4": (30000 / 130 * (1 + 3e-6), None, 2311, 0, "past"),
    "cfg4f": (30000 / 130 * (1 + 3e-6), None, 2311, 0, "future"),
}


def taps_of(name):
    period, phw, hw, omit, direction = CASES[name]
    return oracle.tap_offsets(period, period / 50 if phw is None else phw, hw, omit, direction)


def parity():
    rng = np.random.default_rng(1)
    worst = 0.0
    for name in CASES:
        taps = taps_of(name)
        for shape in [(3, 50_000), (1, 7001), (2, 1999), (5, 20_011)]:
            x = rng.standard_normal(shape) * 3 + 10
            want = oracle.apply_filter_direct(x, taps)
            d_x = torch.from_numpy(x).cuda()
            for kern in (K.KERNEL_SPECIALISED, K.KERNEL_GATHER):
                got = eng.filter_device(d_x, taps, kernel=kern).cpu().numpy()
                err = np.abs(got - want).max() / np.abs(x).max()
                worst = max(worst, err)
                flag = "" if err < 1e-12 else "  <-- BAD"
                print(f"{name} {shape} kernel={kern} ({eng.last_filter_kernel}) err={err:.2e}{flag}")
    # odd row stride and odd base offset (TMA alignment path)
    taps = taps_of("cfg2")
    x = rng.standard_normal((4, 30_001))
    buf = torch.zeros(4 * 30_003 + 1, dtype=torch.float64, device="cuda")
    d_x = buf[1:].as_strided((4, 30_001), (30_003, 1))
    d_x.copy_(torch.from_numpy(x))
    got = eng.filter_device(d_x, taps, kernel=K.KERNEL_SPECIALISED).cpu().numpy()
    err = np.abs(got - oracle.apply_filter_direct(x, taps)).max()
    print("odd stride err", err)
    worst = max(worst, err)
    # non-finite samples: window semantics
    x = rng.standard_normal((2, 40_000))
    x[0, 12_345] = np.nan
    x[1, 30_000] = np.inf
    x[1, 5] = 1e12
    want = oracle.apply_filter_direct(x, taps)
    want[~np.isfinite(want)] = 0.0
    d_x = torch.from_numpy(x).cuda()
    for kern in (K.KERNEL_SPECIALISED, K.KERNEL_GATHER):
        got = eng.filter_device(d_x, taps, kernel=kern).cpu().numpy()
        fin = np.isfinite(got).all()
        err = np.abs(got - want).max() / 1e12
        print(f"non-finite kernel={kern} finite={fin} err={err:.2e} zeros={int((got == 0).sum())} "
              f"want zeros={int((want == 0).sum())}")
    return worst


def timing(name="cfg2", n_chans=64, n_samples=1_200_000, tunings=({},)):
    taps = taps_of(name)
    g = torch.Generator(device="cuda").manual_seed(0)
    d_x = torch.randn((n_chans, n_samples), dtype=torch.float64, device="cuda", generator=g)
    d_out = torch.empty_like(d_x)
    rows = []
    for kern, tuning in [(K.KERNEL_GATHER, {})] + [(K.KERNEL_SPECIALISED, t) for t in tunings]:
        try:
            for _ in range(3):
                eng.filter_device(d_x, taps, d_out=d_out, kernel=kern, tuning=tuning)
            torch.cuda.synchronize()
        except RuntimeError as err:
            print(json.dumps(dict(case=name, tuning=tuning, error=str(err)[:120])))
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            eng.filter_device(d_x, taps, d_out=d_out, kernel=kern, tuning=tuning)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gbs = 16 * n_chans * n_samples / ms / 1e6
        rows.append(dict(case=name, kernel=eng.last_filter_kernel, tuning=tuning, ms=round(ms, 4),
                         gbs=round(gbs, 1), frac=round(gbs / 6549.1, 4)))
        print(json.dumps(rows[-1]))
    return rows


if __name__ == "__main__":
    t0 = time.time()
    if "--no-parity" not in sys.argv:
        print("worst", parity())
    def grid(us, pfs=(2, 4), ctas=(1, 2)):
        return [{}] + [{"steps_per_chunk": u, "prefetch_chunks": pf, "ctas_per_sm": c}
                       for u in us for pf in pfs for c in ctas]

    rows = timing("cfg2", 64, 1_200_000, [{}])
    rows += timing("cfg3", 64, 1_200_000, [{}, {"prefetch_chunks": 3}, {"prefetch_chunks": 4}])
    rows += timing("cfg4", 64, 1_200_000, [{}, {"steps_per_chunk": 8, "prefetch_chunks": 2, "ctas_per_sm": 1}])
    rows += timing("cfg4f", 64, 1_200_000, [{}])
    rows += timing("cfg1", 64, 1_200_000, [{}])
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/check_comb_e.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    print("done in", round(time.time() - t0, 1), "s")
