"""Multi-GPU data parallelism for stacks: one process per GPU, contiguous slice ranges per rank,
and ONE small collective — an all-gather of the per-slice result rows (metrics, validation
scalars, safeguard flags; <= 54 doubles per slice).  Pixels never cross GPUs: every function of
the hot path is per 2-D image (the reference collapses stacks to single slices,
pipeline/dicom_io.py:72-73), so there is no data-path exchange to fuse with a kernel.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def slice_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [a, b) of slices owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def gather_rows(rows: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather row blocks of unequal length (slice_range order) into an [n_total, K] tensor on
    every rank.  Uses all_gather_into_tensor when the blocks are equal (NCCL fast path)."""
    if not dist.is_available() or not dist.is_initialized():
        return rows
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = rows.shape[1]
    sizes = [slice_range(n_total, r, world) for r in range(world)]
    counts = [b - a for a, b in sizes]
    assert rows.shape[0] == counts[rank], (rows.shape, counts, rank)
    if len(set(counts)) == 1:
        out = torch.empty((n_total, k), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(out, rows.contiguous(), group=group)
        return out
    m = max(counts)
    padded = torch.zeros((m, k), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def gather_labels(labels, group: Optional[dist.ProcessGroup] = None):
    """Applied-operation label lists of every rank, concatenated in slice_range order (python
    objects over the process group's object channel; a few bytes per slice, needed only where the
    gathered rows are persisted or reported, pipeline/storage.py)."""
    if not dist.is_available() or not dist.is_initialized():
        return list(labels)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, list(labels), group=group)
    return [l for part in parts for l in part]
