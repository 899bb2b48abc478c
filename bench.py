#!/usr/bin/env python
"""Benchmark of the PARRM hot paths on B200: ``python bench.py --gpus N --steps K --warmup W``.

Headline (BASELINE.json ``metric``): ``filter_data`` channel-samples/s on configs[1] --
synthetic 64-channel LFP, 2 kHz, 130 Hz stimulation artefact, 10 min (64 x 1 200 000 float64),
``create_filter(filter_half_width=2000, filter_direction="both")`` (160 taps).  One *step* is
one pass of the filter over the whole recording.

* ``value``  -- recording resident in HBM, one ``parrm_filter_apply`` launch per step, timed with
  CUDA events on the launching stream (inputs are 614 MB, larger than the 126 MB L2).
* ``e2e``    -- the same pass through the public API ``PARRM.filter_data()`` with a pinned host
  array in and a host array out: H2D + kernel + D2H inside the timed region every step.
* ``roofline`` -- 16 algorithmic bytes per channel-sample (8 read + 8 written) / kernel time,
  against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* ``cpu_baseline`` -- the reference's arithmetic for this path (two SciPy FFT convolutions,
  oracle port) on a bounded sample of the same recording, on this host.
* ``find_period`` -- second metric of BASELINE.json: candidate periods/s of the evaluator on
  the same recording (run-3 shape: ~24.7 k random samples x 64 channels, 20 harmonics, the
  388-candidate grid), with its FP64-pipe roofline and CPU baseline.

``--impl reference`` times the reference's CPU implementation (oracle port; the reference is
pure Python and cannot be pip-installed here: its build backend ``hatchling`` is absent) on
bounded samples of the same workload; rank 0 only.

Multi-GPU (torchrun, one rank per GPU): weak scaling.  Filtering shards by channel with no
collective -- every rank filters its own 64-channel recording; the period search shards the
candidate grid and all-gathers the errors over NCCL.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CHANS, N_SAMPLES, FS, FA = 64, 1_200_000, 2000, 130
HALF_WIDTH = 2000
WORKLOAD = ("cfg2: synthetic 64-ch LFP, 2 kHz, 130 Hz DBS artefact, 10 min (64 x 1.2M f64), "
            "filter_half_width=2000, bidirectional")
BYTES_PER_CHANNEL_SAMPLE = 16.0  # float64: 8 read + 8 written


# ----------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            peaks = json.load(fh)
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_setup(n_gpus: int):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        # before any pinned allocation: keep this rank's host buffers on its GPU's NUMA node
        from pyparrm_b200._sharding import bind_host_to_gpu

        bind_host_to_gpu(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return dist, world, rank, local
    torch.cuda.set_device(0)
    return None, 1, 0, 0


def max_over_ranks(dist, seconds: float) -> float:
    if dist is None:
        return seconds
    import torch

    t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist):
    import torch

    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------- CPU legs
def cpu_filter_pass(oracle, sample: np.ndarray, filt: np.ndarray) -> float:
    t0 = time.perf_counter()
    oracle.apply_filter_fft(sample, filt)
    return time.perf_counter() - t0


def cpu_filter_baseline(data, filt, n_sample_chans=32):
    from oracle import parrm_oracle as oracle

    sample = data[:n_sample_chans]
    seconds = cpu_filter_pass(oracle, sample, filt)
    return {
        "value": sample.size / seconds, "unit": "channel-samples/s", "cores": 1, "kind": "port",
        "sample": f"{n_sample_chans} of {data.shape[0]} channels x {data.shape[1]} samples, 1 pass, "
                  f"{seconds:.1f} s; scipy.signal.convolve (FFT) twice as parrm.py:861-866, "
                  "single-threaded as in the reference",
    }


def cpu_search_baseline(data, indices, periods, bandwidth, n_candidates=None):
    from oracle import parrm_oracle as oracle

    cores = os.cpu_count() or 1
    n_candidates = n_candidates or max(cores, 8)
    z = oracle.standardise(data, 3.0)
    pick = periods[np.linspace(0, len(periods) - 1, n_candidates).astype(int)]
    t0 = time.perf_counter()
    oracle.objective_many(pick, z, indices, bandwidth, 1.0, data.shape[0], n_jobs=cores)
    seconds = time.perf_counter() - t0
    return {
        "value": n_candidates / seconds, "unit": "candidates/s", "cores": cores, "kind": "port",
        "sample": f"{n_candidates} of {len(periods)} grid candidates x {len(indices)} samples x "
                  f"{data.shape[0]} channels, bw={bandwidth}, {cores} threads (pqdm.threads "
                  f"equivalent), {seconds:.1f} s",
    }


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import parrm_oracle as oracle
    from pyparrm_b200.synthetic import make_recording, true_period

    period = true_period(FS, FA)
    filt = oracle.build_filter(period, period / 50, HALF_WIDTH, 0, "both")
    probe = make_recording(2, N_SAMPLES, FS, FA, seed=0)
    per_chan = cpu_filter_pass(oracle, probe, filt) / 2
    budget = 150.0
    n_chans = int(max(1, min(N_CHANS, budget / ((args.steps + args.warmup) * per_chan))))
    sample = make_recording(n_chans, N_SAMPLES, FS, FA, seed=0)
    for _ in range(args.warmup):
        cpu_filter_pass(oracle, sample, filt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.apply_filter_fft(sample, filt)
    seconds = time.perf_counter() - t0
    value = sample.size * args.steps / seconds
    line = {
        "impl": "reference", "metric": "filter_data channel-samples/sec", "value": value,
        "unit": "channel-samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * seconds / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "timing": "host wall clock (CPU path)"},
        "cpu_baseline": {
            "value": value, "unit": "channel-samples/s", "cores": 1, "kind": "port",
            "sample": f"{n_chans} of {N_CHANS} channels x {N_SAMPLES} samples per step; oracle port "
                      "of parrm.py:861-869 (scipy.signal.convolve, FFT, single-threaded as the "
                      "reference runs it); reference not pip-installable here (hatchling absent)",
        },
        "e2e": {"value": value, "unit": "channel-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch

    from oracle import parrm_oracle as oracle  # cpu_baseline leg only
    from pyparrm_b200 import PARRM, _engine, _native, pinned_empty
    from pyparrm_b200.synthetic import make_recording, true_period

    dist, world, rank, local = dist_setup(args.gpus)
    engine = _engine.get_engine()
    hbm_peak, peak_source = measured_peaks()

    # this rank's recording (weak scaling: every GPU gets a full cfg2 recording)
    data = pinned_empty((N_CHANS, N_SAMPLES))
    make_recording(N_CHANS, N_SAMPLES, FS, FA, seed=rank, out=data)
    period = true_period(FS, FA)
    parrm = PARRM(data, FS, FA, verbose=False)
    parrm._period = np.float64(period)  # the filter benchmark does not depend on the search
    parrm.create_filter(filter_half_width=HALF_WIDTH, filter_direction="both")
    taps = (np.flatnonzero(parrm.filter < 0) - HALF_WIDTH).astype(np.int32)
    units = N_CHANS * N_SAMPLES

    d_x = torch.from_numpy(data).cuda()
    d_y = torch.empty_like(d_x)
    stream = torch.cuda.current_stream()

    with ClockSampler(local) as clocks:
        # ---- value: device resident -------------------------------------------------
        for _ in range(args.warmup):
            engine.filter_device(d_x, taps, d_out=d_y)
        barrier(dist)
        launches0 = engine.launches
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        for _ in range(args.steps):
            engine.filter_device(d_x, taps, d_out=d_y)
        stop.record(stream)
        barrier(dist)
        launches = engine.launches - launches0
        dev_seconds = max_over_ranks(dist, start.elapsed_time(stop) * 1e-3)
        kernel_seconds = start.elapsed_time(stop) * 1e-3 / args.steps  # one launch per step

        # ---- e2e: public API, host in / host out -----------------------------------
        for _ in range(min(args.warmup, 3)):
            parrm.filter_data()
        e2e_steps = max(3, min(args.steps, 20))
        barrier(dist)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = parrm.filter_data()
        torch.cuda.synchronize()
        e2e_seconds = max_over_ranks(dist, time.perf_counter() - t0)
        barrier(dist)

        # ---- find_period evaluator (second metric) ---------------------------------
        search = bench_search(engine, dist, world, rank, data, args)

    # parity guard on what was just timed (3 channels against the oracle's direct form)
    want = oracle.apply_filter_direct(data[:3], taps)
    parity = float(np.abs(out[:3] - want).max() / np.abs(data[:3]).max())
    assert parity <= 1e-9, f"bench parity check failed: {parity:.3e}"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    filt = oracle.build_filter(period, period / 50, HALF_WIDTH, 0, "both")
    cpu = cpu_filter_baseline(data, filt)
    achieved = BYTES_PER_CHANNEL_SAMPLE * units / kernel_seconds / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "filter_traffic.json")) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": "filter_data channel-samples/sec",
        "value": world * units * args.steps / dev_seconds,
        "unit": "channel-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_seconds / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "taps": int(taps.shape[0]),
            "sharding": "one 64-channel recording per GPU, no collective" if world > 1 else "single GPU",
            "l2": "inputs (614 MB/GPU) larger than L2 (126 MB); no flush needed",
            "timing": "CUDA events on the launching stream, max over ranks",
        },
        "clocks": clocks.summary(),
        "e2e": {
            "value": world * units * e2e_steps / e2e_seconds, "unit": "channel-samples/s",
            "h2d_bytes_per_step": units * 8, "d2h_bytes_per_step": units * 8,
            "steps": e2e_steps, "ms_per_step": 1e3 * e2e_seconds / e2e_steps,
            "api": "PARRM.filter_data() on a pinned NumPy array, NumPy result",
        },
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic,
            "kernel": "parrm_filter_apply", "peak_source": peak_source,
            "algorithmic_bytes_per_launch": BYTES_PER_CHANNEL_SAMPLE * units,
        },
        "cpu_baseline": cpu,
        "parity_max_rel_err": parity,
        "find_period": search,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def bench_search(engine, dist, world, rank, data, args):
    """Evaluator throughput on the run-3 shape of this recording; candidates sharded over ranks."""
    import torch

    from pyparrm_b200 import _native

    n_chans, n_samples = data.shape
    rng = np.random.default_rng(0)
    lo, hi = int(np.floor(0.025 * n_samples)), int(n_samples - 2 - np.ceil(0.025 * n_samples))
    indices = np.unique(rng.integers(0, hi - lo, 25000)) + lo       # parrm.py:359-374
    bandwidth = 20
    p0 = FS / FA
    grid = np.unique(p0 * np.concatenate((1 + np.arange(-1e-2, 1e-2 + 1e-4, 1e-4) / 3,
                                          1 + np.arange(-1e-3, 1e-3 + 1e-5, 1e-5) / 3)))
    per_rank = 8 * len(grid)                                          # weak scaling: fixed per GPU
    periods = np.resize(grid, per_rank) * (1 + 1e-9 * rank)
    (tile,) = engine.prepare_tiles(data, [indices], 3.0)
    d_periods = torch.from_numpy(periods).cuda()
    stream = torch.cuda.current_stream()
    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        d_err = engine.evaluate_device(tile, d_periods, bandwidth, 1.0, n_chans)
    barrier(dist)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for _ in range(steps):
        d_err = engine.evaluate_device(tile, d_periods, bandwidth, 1.0, n_chans)
        if dist is not None:  # the one exchange step of the sharded search
            gathered = torch.empty(world * per_rank, dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(gathered, d_err)
    stop.record(stream)
    barrier(dist)
    seconds = max_over_ranks(dist, start.elapsed_time(stop) * 1e-3)

    # FP64 FMA peak of this GPU, measured (no figure for it in MEASURED_PEAKS.json)
    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    flops = _native.ctypes.c_double(0.0)
    iters = 1 << 15
    for _ in range(2):
        _native.lib.parrm_fp64_fma_burn(iters, sink.data_ptr(), _native.ctypes.byref(flops),
                                        stream.cuda_stream)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    _native.lib.parrm_fp64_fma_burn(iters, sink.data_ptr(), _native.ctypes.byref(flops),
                                    stream.cuda_stream)
    b1.record(stream)
    torch.cuda.synchronize()
    fp64_peak = flops.value / (b0.elapsed_time(b1) * 1e-3) / 1e12

    m = 2 * bandwidth + 1
    n_idx = len(indices)
    # flops the device formulation performs per candidate: right-hand sides
    # 2*N*M*C, harmonic recurrence + sums ~ (6+2)*N*2bw*2, solve (2/3)M^3 + 2*C*M^2
    flops_per_cand = 2.0 * n_idx * m * n_chans + 16.0 * n_idx * 2 * bandwidth * 2 \
        + (2.0 / 3.0) * m ** 3 + 2.0 * n_chans * m * m
    reference_flops_per_cand = n_idx * (m * (m + 1) + 4.0 * n_chans * m)
    achieved = flops_per_cand * per_rank * steps / seconds / 1e12
    result = {
        "metric": "find_period candidates/sec",
        "value": world * per_rank * steps / seconds, "unit": "candidates/s",
        "ms_per_step": 1e3 * seconds / steps, "steps": steps,
        "config": {"workload": f"evaluator on cfg2 run-3 shape: {n_idx} random samples x {n_chans} "
                               f"channels, bandwidth {bandwidth}, {per_rank} candidates per GPU per step",
                   "sharding": "candidates sharded, one NCCL all-gather of errors per step"
                               if world > 1 else "single GPU"},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "traffic": None,
                     "peak_source": "measured here: parrm_fp64_fma_burn (DFMA chains on all SMs)",
                     "flops_per_candidate": flops_per_cand,
                     "reference_flops_per_candidate": reference_flops_per_cand},
    }
    # the whole search through the public API (host array in, period out): three coarse-to-fine
    # grid runs + lock-step Nelder-Mead, ~1 800 objective evaluations
    from pyparrm_b200 import PARRM
    from pyparrm_b200.synthetic import true_period

    searcher = PARRM(data, FS, FA, verbose=False)
    launches0 = engine.launches
    t0 = time.perf_counter()
    searcher.find_period(random_seed=0)
    api_seconds = time.perf_counter() - t0
    result["public_api"] = {
        "call": "PARRM(data, 2000, 130).find_period(random_seed=0) on the 64 x 1.2M recording",
        "seconds": api_seconds, "gpu_launches": engine.launches - launches0,
        "period": float(searcher.period),
        "rel_err_vs_injected_period": abs(float(searcher.period) - true_period(FS, FA))
        / true_period(FS, FA),
    }
    if rank == 0:
        result["cpu_baseline"] = cpu_search_baseline(data, indices, grid, bandwidth)
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
