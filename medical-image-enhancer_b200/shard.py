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


# ------------------------------------------------------------------------------------------
# Cohort entry point: shard -> per-rank stack pipeline -> device-resident row gather -> persist
# ------------------------------------------------------------------------------------------
def cohort_spans(sizes, rank: int, world: int):
    """Slices of a cohort (volumes of `sizes[v]` slices, concatenated in order) owned by `rank`:
    the contiguous global range of `slice_range`, cut at volume boundaries.
    Returns [(volume, first slice, end slice)], empty volumes skipped."""
    total = int(sum(sizes))
    a, b = slice_range(total, rank, world)
    spans, start = [], 0
    for v, n in enumerate(sizes):
        lo, hi = max(a, start), min(b, start + int(n))
        if lo < hi:
            spans.append((v, lo - start, hi - start))
        start += int(n)
    return spans


class CohortResult:
    """What `process_cohort` returns on every rank.

    local       [(volume, first, end, enhanced, StackResult)] for the spans this rank processed
                (enhanced: device tensor, or host array for host inputs; pixels never leave the rank)
    rows        [total, PACK_COLS] float64 on the device: the gathered per-slice result rows of the
                WHOLE cohort in cohort order (`batch.StackResult` layout), identical on every rank
    counts      slices per rank
    labels      applied-operation labels of the whole cohort (only when gathered)
    run_ids     ids written by the persisting rank (only with `persist=`)"""

    def __init__(self, local, rows, counts, labels=None, run_ids=None):
        self.local, self.rows, self.counts, self.labels, self.run_ids = local, rows, counts, labels, run_ids

    def rows_host(self):
        return self.rows.cpu().numpy()


def _gather_variable(rows: torch.Tensor, counts, group) -> torch.Tensor:
    """All-gather of per-rank row blocks with the given counts (device tensors in, device tensor out)."""
    world = len(counts)
    if world == 1:
        return rows
    k = rows.shape[1]
    if len(set(counts)) == 1:
        out = torch.empty((sum(counts), k), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(out, rows.contiguous(), group=group)
        return out
    m = max(counts)
    padded = torch.zeros((m, k), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def process_cohort(volumes, plan, group: Optional[dist.ProcessGroup] = None, *, already_sharded: bool = False,
                   chunk: Optional[int] = None, workers: int = 2, keep_enhanced: bool = True, ops=None,
                   out_dtype=None, pinned_ins=None, pinned_outs=None, schedule=None,
                   gather_label_lists: bool = False, persist: Optional[dict] = None, processor=None) -> CohortResult:
    """The hot path over a cohort of volumes on all ranks of `group` (one process per GPU).

    volumes      sequence of [N_v, H, W] stacks: CUDA tensors (uint16 bit pattern in int16 / uint16,
                 or float32) -> the device-resident pipeline (`batch.process_stack`), or host numpy
                 arrays -> the host-buffer pipeline with overlapped copies (`batch.process_stacks_host`).
                 By default every rank is handed the WHOLE cohort (e.g. memory-mapped) and takes its
                 contiguous slice range (`cohort_spans`); with `already_sharded=True` the sequence is
                 this rank's own shard (inputs loaded per rank) and only the counts are exchanged.
    The one collective: an all-gather of the per-slice result rows, sent straight from the device buffer
    the pipeline wrote them to (`StackResult.packed_dev`) -- `all_gather_into_tensor` over NCCL when the
    shards are equal.  Enhanced pixels stay on the rank that produced them.
    persist      optional dict(input_filename=..., plan_json=..., metadata_summary=...): rank 0 writes one
                 `runs` record per slice in one transaction (`pipeline.storage.save_stack`); implies the
                 label gather.
    processor    test hook: callable(stack, plan) -> (enhanced, StackResult) replacing the CUDA pipeline."""
    import numpy as np

    from . import batch

    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    volumes = list(volumes)
    sizes = [int(v.shape[0]) for v in volumes]
    spans = [(v, 0, n) for v, n in enumerate(sizes) if n] if already_sharded else cohort_spans(sizes, rank, world)
    parts = [volumes[v][a:b] for v, a, b in spans]

    local = []
    if processor is not None:
        for (v, a, b), part in zip(spans, parts):
            enh, res = processor(part, plan)
            local.append((v, a, b, enh, res))
    elif parts and isinstance(parts[0], np.ndarray):
        pins = [pinned_ins[v][a:b] if pinned_ins is not None and pinned_ins[v] is not None else None for v, a, b in spans]
        pouts = [pinned_outs[v][a:b] if pinned_outs is not None and pinned_outs[v] is not None else None for v, a, b in spans]
        results = batch.process_stacks_host(parts, plan, chunk=chunk, ops=ops, workers=workers, schedule=schedule,
                                            pinned_ins=pins, pinned_outs=pouts,
                                            out_dtype=out_dtype if out_dtype is not None else np.float32)
        for (v, a, b), (enh, res) in zip(spans, results):
            local.append((v, a, b, enh, res))
    else:
        for (v, a, b), part in zip(spans, parts):
            res = batch.process_stack(part, plan, chunk=chunk, keep_enhanced=keep_enhanced, ops=ops, workers=workers)
            local.append((v, a, b, res.enhanced, res))

    # ---- the collective: device-resident rows -> all ranks ----
    blocks = [res.packed_dev if res.packed_dev is not None else torch.from_numpy(res.packed) for *_, res in local]
    if blocks:
        rows = blocks[0] if len(blocks) == 1 else torch.cat(blocks, dim=0)
    else:
        dev = ops.device if ops is not None else (torch.device("cuda", torch.cuda.current_device())
                                                  if torch.cuda.is_available() and processor is None else torch.device("cpu"))
        rows = torch.empty((0, batch.PACK_COLS), dtype=torch.float64, device=dev)
    n_local = int(rows.shape[0])
    if distributed and world > 1:
        if already_sharded:
            cnt = torch.tensor([n_local], dtype=torch.int64, device=rows.device)
            allc = [torch.empty_like(cnt) for _ in range(world)]
            dist.all_gather(allc, cnt, group=group)
            counts = [int(c.item()) for c in allc]
        else:
            total = sum(sizes)
            counts = [b - a for a, b in (slice_range(total, r, world) for r in range(world))]
        gathered = _gather_variable(rows, counts, group)
    else:
        counts, gathered = [n_local], rows

    labels = None
    if gather_label_lists or persist is not None:
        labels = gather_labels([lab for *_, res in local for lab in res.labels], group)
    run_ids = None
    if persist is not None and rank == 0:
        from .pipeline import storage
        run_ids = storage.save_stack(gathered.cpu().numpy(), labels, persist.get("input_filename", "cohort"),
                                     plan_json=persist.get("plan_json", ""),
                                     metadata_summary=persist.get("metadata_summary"))
    return CohortResult(local, gathered, counts, labels, run_ids)
