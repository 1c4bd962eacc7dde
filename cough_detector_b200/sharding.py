"""Clip / stream sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every reduction on the hot path is per clip, so ranks never exchange data while
computing; the only collective is the final gather of the (B, 2) logits (8 bytes per clip).
Streams shard by time: each rank gets a contiguous run of windows and the samples those windows
cover (neighbouring shards overlap by window - hop samples), so every window is computed by
exactly one rank and the gathered probabilities are in time order.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

WINDOW = 16000


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to `device_index` (the socket its PCIe root
    hangs off), so the pinned host buffers it allocates afterwards live on that NUMA node.  The host-buffer path is
    PCIe-bound; with one process per GPU and no binding, ranks end up pulling their audio across the socket link.
    Returns False (and changes nothing) when NVML or the topology is not available."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def stream_windows(n_samples: int, hop: int) -> int:
    return 0 if n_samples < WINDOW or hop <= 0 else (n_samples - WINDOW) // hop + 1


def shard_stream(n_samples: int, hop: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """(first_window, n_windows, sample_start, sample_end) of this rank's time shard."""
    w0, w1 = shard_range(stream_windows(n_samples, hop), rank, world)
    if w1 == w0:
        return w0, 0, 0, 0
    return w0, w1 - w0, w0 * hop, (w1 - 1) * hop + WINDOW


def gather_rows(local: torch.Tensor, counts: List[int]) -> torch.Tensor:
    """All-gather per-rank row blocks of unequal length into rank order.  Works on NCCL (cuda tensors)
    and gloo (cpu tensors); pads to the longest block because all_gather wants equal shapes."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
