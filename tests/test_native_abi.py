"""The C-ABI library loads on a CPU-only host and exports what include/parrm_b200.h declares."""

import ctypes
import os
import re

import numpy as np
import pytest

from pyparrm_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "parrm_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(parrm_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 15
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature in _native.py"
    assert set(_native.SIGNATURES) == set(names)


def test_versions_and_constants_agree_with_header():
    text = open(HEADER).read()
    assert int(re.search(r"#define PARRM_B200_ABI_VERSION (\d+)", text).group(1)) == _native.ABI_VERSION
    assert int(re.search(r"#define PARRM_MAX_BANDWIDTH (\d+)", text).group(1)) == _native.MAX_BANDWIDTH
    assert _native.lib.parrm_abi_version() == _native.ABI_VERSION


def test_argument_errors_are_reported_not_crashed():
    status = _native.lib.parrm_build_taps(-1.0, 0.1, 10, 0, 0, None, None, None)
    assert status == 1 and "period" in _native.last_error()
    taps = np.array([3, 2, 1], dtype=np.int32)
    plan = np.zeros(4096, dtype=np.uint8)
    status = _native.lib.parrm_filter_plan(taps.ctypes.data, 3, 0, 0, plan.ctypes.data, 4096)
    assert status == 1 and "ascending" in _native.last_error()
    taps = np.arange(1, 161, dtype=np.int32)
    assert _native.lib.parrm_filter_plan_bytes(taps.ctypes.data, 160) >= 128 + 160 * 4


def test_launch_accounting_follows_the_shape():
    """parrm_eval_launch_count is host arithmetic: accumulate + solve, the column-sum pair of
    the tensor path, the dense copy of y for wide / odd layouts, the partial sums of split
    launches."""
    count = _native.lib.parrm_eval_launch_count
    n = 24_000
    assert count(None, 64, 64, n, 3048, 20) == 4      # one dense tile, one CTA per candidate
    assert count(None, 64, 64, n, 5, 20) == 5         # few candidates: sample splits
    assert count(None, 256, 256, n, 3048, 20) == 5    # four channel tiles: re-tiled copy
    assert count(None, 7, 7, n, 3048, 20) == 5        # odd width: re-tiled copy
    assert count(None, 80, 64, n, 3048, 20) == 5      # padded rows: re-tiled copy
    assert count(None, 1, 1, n, 100_000, 20) == 2     # narrow kernel
    assert count(None, 1, 1, n, 3, 20) == 3
    assert count(None, 64, 64, n, 0, 20) == 0
    assert count(None, 64, 64, n, 10, 99) == 0


def test_no_silent_cpu_path():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyparrm_b200 import PARRM, _engine

    _engine.set_engine(None)
    p = PARRM(np.random.default_rng(0).standard_normal((1, 200)), 20, 10, verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.find_period()
    assert _native.device_count() == 0


@pytest.mark.filterwarnings("ignore::DeprecationWarning")  # fork() in a threaded process: the point
def test_host_copy_is_exact_for_any_size_alignment_and_thread_count():
    """``parrm_host_copy`` (the staged leg of the host pipeline for pageable arrays): every
    byte copied, nothing outside the range touched, for odd sizes, unaligned ends, more threads
    than pieces, and two callers at once (the pool serialises them)."""
    import ctypes
    import threading

    lib = _native.lib
    rng = np.random.default_rng(0)
    for n in (0, 1, 63, 65, 4097, (1 << 20) + 17, 3 * (1 << 20) + 5, 9_600_007):
        for off_s, off_d in ((0, 0), (1, 3), (8, 16)):
            src = rng.integers(0, 255, n + off_s + 64, dtype=np.uint8)
            for threads in (1, 3, 8, 16):
                dst = np.full(n + off_d + 64, 7, dtype=np.uint8)
                assert lib.parrm_host_copy(ctypes.c_void_p(dst.ctypes.data + off_d),
                                           ctypes.c_void_p(src.ctypes.data + off_s), n, threads) == 0
                assert np.array_equal(dst[off_d:off_d + n], src[off_s:off_s + n]), (n, threads)
                assert (dst[:off_d] == 7).all() and (dst[off_d + n:] == 7).all()
    assert lib.parrm_host_copy(None, None, 5, 2) != 0 and "null" in _native.last_error()
    assert lib.parrm_host_copy(None, None, 0, 0) != 0  # thread count out of range

    a, b = (rng.integers(0, 255, 6 << 20, dtype=np.uint8) for _ in range(2))
    out = [np.zeros_like(a), np.zeros_like(b)]

    def worker(i, src):
        for _ in range(5):
            out[i][:] = 0
            assert lib.parrm_host_copy(ctypes.c_void_p(out[i].ctypes.data),
                                       ctypes.c_void_p(src.ctypes.data), src.nbytes, 4) == 0
            assert np.array_equal(out[i], src)

    pool = [threading.Thread(target=worker, args=(i, s)) for i, s in enumerate((a, b))]
    for t in pool:
        t.start()
    for t in pool:
        t.join()

    # a forked child has none of the parent's copy threads: it must start a pool of its own
    pid = os.fork()
    if pid == 0:
        import signal

        signal.alarm(30)  # a deadlocked child must not hang the suite
        out[0][:] = 0
        rc = lib.parrm_host_copy(ctypes.c_void_p(out[0].ctypes.data), ctypes.c_void_p(a.ctypes.data),
                                 a.nbytes, 4)
        os._exit(0 if rc == 0 and np.array_equal(out[0], a) else 1)
    assert os.WEXITSTATUS(os.waitpid(pid, 0)[1]) == 0
