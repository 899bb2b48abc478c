"""Sharding of the two hot paths over the ranks of a ``torch.distributed`` group.

One process per GPU:

* ``filter_data`` partitions by channel -- channels are independent, so there is **no
  collective** -- or, when there are fewer channels than ranks, by time with halos of the
  tap window read from the source recording (``parrm_filter_apply``'s ``x_t0 / t0 / n_out``).
* ``find_period`` partitions the candidate grid of every run into contiguous blocks; the one
  exchange step per run is an all-gather of the ``P`` float64 fit errors (8 P bytes), after
  which every rank ranks the full grid exactly as the reference does (``parrm.py:456-463``).
  The <= 25-point Nelder-Mead rounds are evaluated on every rank (they are deterministic), so
  all ranks finish with the bit-identical period.

The functions work with any backend: tensors live on the GPU for ``nccl`` and on the host for
``gloo`` (the CPU tests run world size 2 over gloo with a stand-in engine).
"""

from __future__ import annotations

import numpy as np

_group = None
_enabled = False


def enable(group=None) -> None:
    """Shard ``find_period`` / ``filter_sharded`` over ``group`` (default: the world group)."""
    global _group, _enabled
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable() -> None:
    global _group, _enabled
    _group, _enabled = None, False


def active() -> bool:
    return _enabled


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


def bind_host_to_gpu(device_index: int) -> list[int]:
    """Pin this process to the CPUs local to a GPU's PCIe root before it allocates pinned memory.

    With one process per GPU every rank streams its shard through its own pinned ring
    (``_engine.filter_host``); first-touch then places the ring on the GPU's NUMA node instead
    of wherever the launcher happened to start the process.  Returns the CPU list used ([] when
    the topology is not exposed -- nothing is changed then).
    """
    import os

    address = _pci_address(device_index)
    if address is None:
        return []
    try:
        with open(f"/sys/bus/pci/devices/{address}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
    except (OSError, ValueError):
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


def evaluate_sharded(evaluate, periods: np.ndarray) -> np.ndarray:
    """Fit errors of all ``periods``: this rank evaluates its block, one all-gather exchanges
    the blocks.  ``evaluate(block_of_periods) -> float64 array``; identical result on every
    rank."""
    import torch
    import torch.distributed as dist

    periods = np.ascontiguousarray(periods, dtype=np.float64).ravel()
    world, rank = _world_rank()
    n = periods.shape[0]
    if world == 1 or n == 0:
        return evaluate(periods)
    per = -(-n // world)
    lo, hi = block(n, world, rank)
    mine = np.full(per, np.nan)
    if hi > lo:
        mine[: hi - lo] = evaluate(periods[lo:hi])
    on_gpu = dist.get_backend(_group) == "nccl"
    send = torch.from_numpy(mine)
    if on_gpu:
        send = send.cuda()
    recv = torch.empty(world * per, dtype=torch.float64, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=_group)
    return recv.cpu().numpy()[:n].copy()


def filter_sharded(engine, data: np.ndarray, taps: np.ndarray):
    """This rank's share of ``filter_data``: returns ``(out, (c0, c1, t0, t1))`` with
    ``out[c - c0, t - t0]`` the filtered samples.  No communication."""
    world, rank = _world_rank()
    taps = np.asarray(taps)
    w_lo, w_hi = min(int(taps[0]), 0), max(int(taps[-1]), 0)
    n_chans, n_samples = data.shape
    c0, c1, t0, t1, x0, x1 = channel_or_time_shards(n_chans, n_samples, world, w_lo, w_hi)[rank]
    if c1 <= c0 or t1 <= t0:
        return np.empty((0, 0), dtype=np.float64), (c0, c1, t0, t1)
    if (t0, t1) == (0, n_samples):
        return engine.filter_host(data[c0:c1], taps), (c0, c1, t0, t1)
    out = engine.filter_host_window(data[c0:c1, x0:x1], taps, x0, t0, t1, n_samples)
    return out, (c0, c1, t0, t1)
