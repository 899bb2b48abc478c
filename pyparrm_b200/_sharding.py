"""Sharding of the two hot paths over the ranks of a ``torch.distributed`` group.

One process per GPU:

* ``filter_data`` partitions by channel -- channels are independent, so there is **no
  collective** -- or, when there are fewer channels than ranks, by time with halos of the
  tap window read from the source recording (``parrm_filter_apply``'s ``x_t0 / t0 / n_out``).
* ``find_period`` standardises by channel block (each rank uploads only its rows) and
  replicates the small search tiles with one all-gather each; it then partitions the
  candidate grid of every run into contiguous blocks; the one exchange step per run is an
  all-gather of the ``P`` float64 fit errors (8 P bytes, on the device for nccl), after which
  every rank ranks the full grid exactly as the reference does (``parrm.py:456-463``).
  Sweeps that only need the winner exchange one ``(error, index)`` pair per rank
  (:func:`minloc_sharded`).
  The <= 25-point Nelder-Mead rounds are evaluated on every rank (they are deterministic), so
  all ranks finish with the bit-identical period.

The functions work with any backend: tensors live on the GPU for ``nccl`` and on the host for
``gloo`` (the CPU tests run world size 2 over gloo with a stand-in engine).
"""

from __future__ import annotations

import numpy as np

_group = None
_enabled = False
_gather = "rank0"
GATHER_MODES = ("rank0", "all", "none")


def enable(group=None, gather: str = "rank0", device: int | None = None) -> None:
    """Shard ``PARRM.find_period`` and ``PARRM.filter_data`` over ``group`` (default: the
    world group).  Every rank calls the same methods on the same recording (SPMD).

    ``gather`` says what ``filter_data`` returns: ``"rank0"`` -- the stitched ``[C, T]`` result
    on rank 0, the rank's own shard elsewhere; ``"all"`` -- the stitched result on every rank;
    ``"none"`` -- every rank keeps only its shard (``PARRM.filter_shard`` says which), no
    communication at all: the scalable mode, since the result never crosses NVLink or PCIe
    twice.  ``find_period`` returns the bit-identical period on every rank in all modes.

    With the ``nccl`` backend the rank's GPU is ``device`` (default: ``LOCAL_RANK``, else the
    current device); it is made current before the engine is created, and an engine already
    bound to another GPU is an error rather than a silent pile-up on GPU 0.
    """
    global _group, _enabled, _gather
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if gather not in GATHER_MODES:
        raise ValueError(f"`gather` must be one of {GATHER_MODES}.")
    if dist.get_backend(group) == "nccl":
        import os

        import torch

        from . import _engine

        if device is None:
            device = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
        torch.cuda.set_device(device)
        engine = _engine.current_engine()
        bound = getattr(getattr(engine, "device", None), "index", None)
        if bound is not None and bound != device:
            raise RuntimeError(
                f"the engine of this process is bound to cuda:{bound} but this rank's GPU is "
                f"cuda:{device}; call torch.cuda.set_device(LOCAL_RANK) before the first "
                "pyparrm_b200 call, or pass `device=` to enable_sharding()")
    _group, _enabled, _gather = group, True, gather


def disable() -> None:
    global _group, _enabled, _gather
    _group, _enabled, _gather = None, False, "rank0"


def active() -> bool:
    return _enabled


def gather_mode() -> str:
    return _gather


def _parse_cpulist(text: str) -> list[int]:
    """Expand a sysfs CPU list such as ``0-15,64-79``."""
    cpus: list[int] = []
    for part in filter(None, text.strip().split(",")):
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def _pci_address(device_index: int) -> str | None:
    """``dddd:bb:dd.f`` of a CUDA device as sysfs spells it, or None."""
    import torch

    props = torch.cuda.get_device_properties(device_index)
    if all(hasattr(props, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
        return f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    try:  # older torch: ask NVML by UUID (robust to CUDA_VISIBLE_DEVICES)
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{props.uuid}".encode())
        bus_id = pynvml.nvmlDeviceGetPciInfo(handle).busId
        bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
        domain, rest = bus_id.lower().split(":", 1)
        return f"{domain[-4:]}:{rest}"
    except Exception:  # noqa: BLE001 - topology is optional
        return None


def _nvml_cpu_affinity(device_index: int) -> list[int]:
    """CPUs NVML reports as local to the GPU (the "CPU Affinity" column of nvidia-smi topo)."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        handle = pynvml.nvmlDeviceGetHandleByUUID(f"GPU-{props.uuid}".encode())
        n_words = (os_cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        return [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
    except Exception:  # noqa: BLE001 - topology is optional
        return []


def os_cpu_count() -> int:
    import os

    return os.cpu_count() or 1


def bind_host_to_gpu(device_index: int) -> list[int]:
    """Pin this process to the CPUs local to a GPU's PCIe root before it allocates pinned memory.

    With one process per GPU every rank streams its shard through its own pinned ring
    (``_engine.filter_host``); first-touch then places the ring on the GPU's NUMA node instead
    of wherever the launcher happened to start the process.  Returns the CPU list used ([] when
    the topology is not exposed -- nothing is changed then).
    """
    import os

    cpus: list[int] = []
    address = _pci_address(device_index)
    if address is not None:
        try:
            with open(f"/sys/bus/pci/devices/{address}/local_cpulist") as f:
                cpus = _parse_cpulist(f.read())
        except (OSError, ValueError):
            cpus = []
    if not cpus:  # containers often hide sysfs topology: ask the driver (NVML) instead
        cpus = _nvml_cpu_affinity(device_index)
    if not cpus:
        return []
    allowed = sorted(set(cpus) & os.sched_getaffinity(0))
    if not allowed:
        return []
    try:
        os.sched_setaffinity(0, allowed)
    except OSError:
        return []
    return allowed


def _world_rank():
    import torch.distributed as dist

    return dist.get_world_size(_group), dist.get_rank(_group)


def block(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``n`` items owned by ``rank`` (equal sizes, the last
    blocks may be short or empty)."""
    per = -(-n // world) if n else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def channel_or_time_shards(n_chans: int, n_samples: int, world: int, w_lo: int, w_hi: int):
    """Per-rank work for the filter: ``(c0, c1, t0, t1, x0, x1)`` -- channels ``[c0, c1)``,
    outputs ``[t0, t1)``, input samples ``[x0, x1)`` (outputs plus the tap-window halo)."""
    if n_chans >= world or n_samples == 0:
        return [(*block(n_chans, world, r), 0, n_samples, 0, n_samples) for r in range(world)]
    # fewer channels than ranks: every channel is cut in time across ranks_per_chan ranks
    shards = []
    ranks_per_chan = world // n_chans
    for r in range(world):
        chan, part = divmod(r, ranks_per_chan)
        if chan >= n_chans:
            shards.append((0, 0, 0, 0, 0, 0))
            continue
        t0, t1 = block(n_samples, ranks_per_chan, part)
        shards.append((chan, chan + 1, t0, t1, max(0, t0 - w_hi), min(n_samples, t1 - w_lo)))
    return shards


def _on_gpu() -> bool:
    import torch.distributed as dist

    return dist.get_backend(_group) == "nccl"


def _all_gather(tensor):
    """``[world, *tensor.shape]`` of a contiguous tensor that has the same shape on every rank
    (device tensor for nccl, host tensor for gloo)."""
    import torch
    import torch.distributed as dist

    world, _ = _world_rank()
    tensor = tensor.contiguous()
    out = torch.empty(world * tensor.numel(), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.reshape(-1), group=_group)
    return out.view((world,) + tuple(tensor.shape))


def _as_comm_tensor(values):
    """NumPy array or tensor -> tensor where the backend wants it (no host bounce for CUDA)."""
    import torch

    if isinstance(values, np.ndarray):
        values = torch.from_numpy(np.ascontiguousarray(values))
    if _on_gpu() and not values.is_cuda:
        values = values.cuda()
    return values


def evaluate_sharded(evaluate, periods: np.ndarray) -> np.ndarray:
    """Fit errors of all ``periods``: this rank evaluates its contiguous block, one all-gather
    exchanges the blocks (8 bytes per candidate), every rank gets the identical full vector --
    which it then ranks exactly as the reference does (parrm.py:456-463).

    ``evaluate(block_of_periods)`` returns a float64 NumPy array or a torch tensor; a CUDA
    tensor stays on the device through the collective (one D2H of the gathered vector)."""
    import torch

    periods = np.ascontiguousarray(periods, dtype=np.float64).ravel()
    world, rank = _world_rank()
    n = periods.shape[0]
    if world == 1 or n == 0:
        mine = evaluate(periods)
        return mine.cpu().numpy() if isinstance(mine, torch.Tensor) else mine
    per = -(-n // world)
    lo, hi = block(n, world, rank)
    if hi > lo:
        mine = _as_comm_tensor(evaluate(periods[lo:hi])).to(torch.float64)
        send = torch.full((per,), float("nan"), dtype=torch.float64, device=mine.device)
        send[: hi - lo] = mine
    else:
        send = _as_comm_tensor(np.full(per, np.nan))
    return _all_gather(send).reshape(-1)[:n].cpu().numpy().copy()


def minloc_sharded(evaluate, periods: np.ndarray) -> tuple[int, float]:
    """Winner of a candidate sweep: ``(index into periods, its fit error)``, identical on every
    rank, NaNs skipped, ties to the lowest index.  Each rank reduces its block to one
    ``(error, global index)`` pair and ONE collective exchanges the pairs -- 16 bytes per rank.
    NCCL has no MINLOC operator and 64 + 32 bits do not pack into one reducible word without
    losing error bits, so the pair travels as two float64 (indices < 2^53 are exact) through
    an all-gather, which costs the same single latency as an all-reduce of this size."""
    import torch

    periods = np.ascontiguousarray(periods, dtype=np.float64).ravel()
    world, rank = _world_rank()
    n = periods.shape[0]
    lo, hi = block(n, world, rank) if world > 1 else (0, n)
    pair = np.array([np.inf, -1.0])
    if hi > lo:
        mine = evaluate(periods[lo:hi])
        if isinstance(mine, torch.Tensor):
            mine = torch.where(torch.isnan(mine), torch.full_like(mine, float("inf")), mine)
            value, index = torch.min(mine, dim=0)
            pair_t = torch.stack([value, (index + lo).to(torch.float64)])
        else:
            mine = np.where(np.isnan(mine), np.inf, mine)
            pair_t = None
            pair = np.array([mine.min(), float(lo + int(mine.argmin()))])
    else:
        pair_t = None
    if world == 1:
        if pair_t is not None:
            pair = pair_t.cpu().numpy()
        return int(pair[1]), float(pair[0])
    send = pair_t if pair_t is not None else _as_comm_tensor(pair)
    pairs = _all_gather(_as_comm_tensor(send)).cpu().numpy()
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))  # by error, then by index
    best = pairs[order[0]]
    return int(best[1]), float(best[0])


def prepare_tiles_sharded(engine, data: np.ndarray, index_sets, outlier_boundary: float,
                          precision: str = "fp64"):
    """Standardised search tiles with the recording read once across ALL ranks: every rank
    standardises its channel block (its rows are the only ones that cross PCIe) and one
    all-gather per tile replicates the ``[samples, channels]`` tiles (<= 51 MB) over NVLink
    (SURVEY.md 8(e)).  Channels are independent in ``_standardise_data`` (parrm.py:272-280),
    so the gathered tile is bit-identical to the unsharded one."""
    world, rank = _world_rank()
    n_chans = data.shape[0]
    if world == 1 or n_chans < world:
        return engine.prepare_tiles(data, index_sets, outlier_boundary, precision)
    per = -(-n_chans // world)
    c0, c1 = block(n_chans, world, rank)
    local = engine.prepare_tiles(data[c0:c1], index_sets, outlier_boundary, precision)
    return [engine.merge_channel_tiles(tile, _all_gather, n_chans, per) for tile in local]


def filter_sharded(engine, data: np.ndarray, taps: np.ndarray, precision: str = "fp64",
                   gather: str = "none", out_dtype=None):
    """This rank's share of ``filter_data``: returns ``(out, (c0, c1, t0, t1), stitched)``.

    ``gather == "none"``: ``out[c - c0, t - t0]`` are the rank's filtered samples, produced by
    the pipelined host path; no communication.  Otherwise the shards are exchanged on the
    device (one all-gather of equal, zero-padded blocks over NVLink) before anything returns
    to the host, and ``out`` is the stitched ``[C, T]`` array on rank 0 (``"rank0"``) or on
    every rank (``"all"``); ``stitched`` says which of the two ``out`` is."""
    world, rank = _world_rank()
    taps = np.asarray(taps)
    w_lo, w_hi = min(int(taps[0]), 0), max(int(taps[-1]), 0)
    n_chans, n_samples = data.shape
    shards = channel_or_time_shards(n_chans, n_samples, world, w_lo, w_hi)
    c0, c1, t0, t1, x0, x1 = shards[rank]
    mine = (c0, c1, t0, t1)
    whole_rows = (t0, t1) == (0, n_samples)
    if gather == "none" or world == 1:
        if c1 <= c0 or t1 <= t0:
            return np.empty((0, 0), dtype=np.float64), mine, False
        if whole_rows:
            if out_dtype is None:
                return engine.filter_host(data[c0:c1], taps, precision), mine, world == 1
            out = engine.filter_host(data[c0:c1], taps, precision, out_dtype=out_dtype)
            return out, mine, world == 1
        out = engine.filter_host_window(data[c0:c1, x0:x1], taps, x0, t0, t1, n_samples)
        return out if out_dtype is None else out.astype(out_dtype), mine, False
    # shards stay on the device, are padded to one common block and all-gathered
    import torch

    rows = max(s[1] - s[0] for s in shards)
    cols = max(s[3] - s[2] for s in shards)
    if c1 > c0 and t1 > t0:
        shard = engine.filter_shard(data[c0:c1, x0:x1], taps, x0, t0, t1, n_samples, precision)
    else:
        shard = np.zeros((0, 0))
    shard = _as_comm_tensor(shard)
    send = torch.zeros((rows, cols), dtype=torch.float64, device=shard.device)
    send[: c1 - c0, : t1 - t0] = shard
    want_full = gather == "all" or rank == 0
    blocks = _all_gather(send)
    cast = (lambda a: a) if out_dtype is None else (lambda a: a.astype(out_dtype))
    if not want_full:
        return cast(shard.cpu().numpy()), mine, False
    if all(s[2:4] == (0, n_samples) for s in shards if s[1] > s[0]) and cols == n_samples:
        full = blocks.reshape(world * rows, cols)[:n_chans]  # channel blocks are contiguous
        return cast(full.cpu().numpy()), mine, True
    out = np.empty((n_chans, n_samples), dtype=np.float64 if out_dtype is None else out_dtype)
    host = blocks.cpu().numpy()
    for r, (sc0, sc1, st0, st1, _, _) in enumerate(shards):
        if sc1 > sc0 and st1 > st0:
            out[sc0:sc1, st0:st1] = host[r, : sc1 - sc0, : st1 - st0]
    return out, mine, True
