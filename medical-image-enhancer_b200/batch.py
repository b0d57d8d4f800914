"""Stack pipeline: normalise -> enhance (plan) -> metrics + validation for [N, H, W] stacks.

This is the batch form of what the reference's runner does per image
(``normalize_image`` -> ``apply_enhancements_from_params`` -> ``compute_metrics(enhanced)`` +
``compute_validation(original, enhanced)``; pipeline/runner.py:80-153, pipeline/tools.py:113-141).
Slices are independent, so a stack is processed in large chunks (default ~128 Mpx): every kernel
launch then covers tens of thousands of CTAs and the per-launch / per-safeguard host round trips
are amortised over hundreds of slices.

Results identical to per-slice calls; ``compute_metrics`` of the same image is evaluated once
per image (the reference recomputes the same values up to four times).
"""

from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .engine import (MC_EDGE_RATIO, MC_NIQE, METRIC_KEYS, Engine, metrics_dict, objective_score,
                     validation_dict)
from .stack import StackOps, get_ops

ROW_COLS = 24
#: packed per-slice result row: metrics_before[24] | metrics_after[24] | ssim psnr | halo noise over | tv_iters | error
PACK_COLS = 2 * ROW_COLS + 2 + 3 + 1 + 1


@dataclass
class StackResult:
    enhanced: Optional[torch.Tensor]        # [N, H, W] float32 on the device (None if not kept)
    packed: np.ndarray                      # [N, PACK_COLS] float64 on the host
    labels: List[List[str]]
    packed_dev: Optional[torch.Tensor] = None   # the same rows, still on the device: the send buffer of the
                                                # multi-GPU row gather (shard.process_cohort), no host round trip

    @property
    def rows_before(self) -> np.ndarray:
        return self.packed[:, :ROW_COLS]

    @property
    def rows_after(self) -> np.ndarray:
        return self.packed[:, ROW_COLS:2 * ROW_COLS]

    def metrics_after(self, i: int) -> Dict[str, float]:
        return metrics_dict(self.rows_after[i])

    def metrics_before(self, i: int) -> Dict[str, float]:
        return metrics_dict(self.rows_before[i])

    def validation(self, i: int) -> Dict[str, object]:
        rb, ra = self.rows_before[i], self.rows_after[i]
        ssim, psnr = self.packed[i, 2 * ROW_COLS], self.packed[i, 2 * ROW_COLS + 1]
        return validation_dict(metrics_dict(rb), metrics_dict(ra), float(ssim), float(psnr),
                               float(rb[MC_NIQE]), float(ra[MC_NIQE]), float(ra[MC_EDGE_RATIO]))

    @property
    def tv_iterations(self) -> np.ndarray:
        return self.packed[:, 2 * ROW_COLS + 5]

    @property
    def failed(self) -> np.ndarray:
        """True where the reference would have raised ValueError for that slice (the slice is
        returned unchanged and its label list holds the error text)."""
        return self.packed[:, 2 * ROW_COLS + 6] != 0

    def score(self, i: int):
        return objective_score(self.validation(i))

    def scores(self) -> np.ndarray:
        return np.array([self.score(i)[0] for i in range(self.packed.shape[0])])


def default_chunk(h: int, w: int, px_budget: int = 1 << 27) -> int:
    """Slices per chunk.  Measured on B200 (profiles/r01_chunk_sweep.txt): the kernels are
    instruction- and latency-bound, not L2-capacity-bound, so large launches win — a chunk is
    sized to ~128 Mpx (512 slices of 512x512, 14 radiographs of 3000x3000), which keeps every grid
    in the tens of thousands of CTAs while leaving several chunks per stack for the host<->device
    copy pipeline of the end-to-end path."""
    return int(max(1, min(4096, px_budget // max(h * w, 1))))


def tapered_schedule(n: int, workers: int = 4, decay: float = 0.78, smallest: int = 16) -> List[int]:
    """Chunk sizes for the end-to-end path: `workers` large chunks first, then geometrically smaller
    ones.  Chunks are claimed dynamically, so the workers finish close together and the copy-out
    that nothing can overlap -- the last chunks' -- is small (measured: 61.9 vs 64.4 ms per
    1024-slice stack against uniform 128-slice chunks)."""
    sizes: List[int] = []
    rem = int(n)
    c = max(smallest, int(round(n / (1.6 * max(1, workers)))))
    k = 0
    while rem > 0:
        take = min(rem, c)
        if rem - take < smallest:          # do not leave a sliver
            take = rem
        sizes.append(take)
        rem -= take
        k += 1
        if k >= workers:
            c = max(smallest, int(c * decay))
    return sizes


def process_chunk(ops: StackOps, raw: torch.Tensor, plan, keep_enhanced: bool = True):
    """One chunk on the current stream.  Returns (enhanced | None, packed device rows, labels)."""
    eng = Engine(ops)
    n = raw.shape[0]
    x = ops.normalize(raw)
    rows_b = ops.metrics(x, with_niqe=True)
    res = eng.enhance_plan(x, plan, rows_before=rows_b, on_error="flag")
    rows_a = res.rows_after
    fr = ops.fullref(x, res.image)
    packed = torch.empty((n, PACK_COLS), dtype=torch.float64, device=ops.device)
    packed[:, :ROW_COLS] = rows_b
    packed[:, ROW_COLS:2 * ROW_COLS] = rows_a
    packed[:, 2 * ROW_COLS:2 * ROW_COLS + 2] = fr
    flags = np.stack([res.halo, res.noise_guard, res.over_processed], axis=1).astype(np.float64)
    packed[:, 2 * ROW_COLS + 2:2 * ROW_COLS + 5] = torch.from_numpy(flags).to(ops.device)
    if res.tv_iterations is not None:
        packed[:, 2 * ROW_COLS + 5] = torch.from_numpy(res.tv_iterations.astype(np.float64)).to(ops.device)
    else:
        packed[:, 2 * ROW_COLS + 5] = 0
    err = np.zeros(n, np.float64)
    err[list(res.errors)] = 1.0
    packed[:, 2 * ROW_COLS + 6] = torch.from_numpy(err).to(ops.device)
    return (res.image if keep_enhanced else None), packed, res.labels


def process_stack(raw: torch.Tensor, plan, chunk: Optional[int] = None, keep_enhanced: bool = True,
                  ops: Optional[StackOps] = None, workers: int = 2) -> StackResult:
    """Device-resident stack (uint16 bit pattern in an int16/uint16 tensor, or float32) ->
    enhanced stack + per-slice metric / validation rows.

    Chunks are independent, so `workers` host threads drive them on separate CUDA streams: the
    small latency-bound kernels, launch gaps and safeguard round trips of one chunk overlap with
    the other chunk's work (the scratch workspace is per thread)."""
    ops = ops or get_ops(raw.device)
    n, h, w = raw.shape
    chunk = chunk or default_chunk(h, w)
    spans = [(a, min(n, a + chunk)) for a in range(0, n, chunk)]
    workers = max(1, min(workers, len(spans)))
    enhanced = torch.empty((n, h, w), dtype=torch.float32, device=ops.device) if keep_enhanced else None
    packed_dev = torch.empty((n, PACK_COLS), dtype=torch.float64, device=ops.device)
    labels: List[Optional[List[List[str]]]] = [None] * len(spans)

    def run_span(i: int) -> None:
        a, b = spans[i]
        enh, packed, lab = process_chunk(ops, raw[a:b], plan, keep_enhanced)
        if keep_enhanced:
            enhanced[a:b] = enh
        packed_dev[a:b] = packed
        labels[i] = lab

    if workers == 1:
        for i in range(len(spans)):
            run_span(i)
    else:
        main = torch.cuda.current_stream(ops.device)
        ready = torch.cuda.Event()
        ready.record(main)
        streams = _worker_streams(ops.device, workers)
        errors: List[BaseException] = []

        def worker(k: int) -> None:
            try:
                with torch.cuda.device(ops.device), torch.cuda.stream(streams[k]):
                    streams[k].wait_event(ready)
                    for i in range(k, len(spans), workers):
                        run_span(i)
            except BaseException as exc:  # noqa: BLE001 - re-raised on the caller's thread
                errors.append(exc)

        threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for st in streams:
            main.wait_stream(st)
        if errors:
            raise errors[0]
    flat = [lab for part in labels for lab in (part or [])]
    return StackResult(enhanced=enhanced, packed=packed_dev.cpu().numpy(), labels=flat, packed_dev=packed_dev)


_stream_cache: dict = {}


def _worker_streams(device: torch.device, count: int):
    key = (device.index, count)
    if key not in _stream_cache:
        _stream_cache[key] = [torch.cuda.Stream(device) for _ in range(count)]
    return _stream_cache[key]


def _spans_of(n: int, chunk: int, schedule: Optional[List[int]]):
    if schedule:
        spans, a, k = [], 0, 0
        while a < n:
            b = min(n, a + max(1, int(schedule[k % len(schedule)])))
            spans.append((a, b))
            a, k = b, k + 1
        return spans
    return [(a, min(n, a + chunk)) for a in range(0, n, chunk)]


def process_stacks_host(raw_hosts: Sequence[np.ndarray], plan, chunk: Optional[int] = None,
                        out_hosts: Optional[Sequence[Optional[np.ndarray]]] = None, ops: Optional[StackOps] = None,
                        pinned_ins: Optional[Sequence[Optional[torch.Tensor]]] = None,
                        pinned_outs: Optional[Sequence[Optional[torch.Tensor]]] = None,
                        workers: int = 2, schedule: Optional[List[int]] = None, out_dtype=np.float32,
                        trace: Optional[list] = None):
    """End-to-end form with HOST buffers for a SEQUENCE of stacks (a cohort of volumes): per chunk,
    host->device copy of the raw slices, the whole pipeline on the GPU, device->host copy of the
    enhanced slices and of the result rows.  `workers` host threads claim chunks dynamically, each
    with its own compute / copy-in / copy-out streams: copies of one chunk overlap the compute of
    the others, and the host round trips inside a chunk (TV live-slice polls, safeguard decisions)
    never leave the GPU idle.  The chunks of all stacks form ONE queue: a worker that finishes the
    last chunk of stack k goes straight on to stack k+1, so the copy-out of a stack's last chunks --
    which nothing inside that stack can overlap -- runs under the next stack's compute; only the
    final stack's tail is exposed.

    raw_hosts[k]: [N, H, W] uint16 or float32 numpy array (ideally backed by pinned memory;
    `pinned_ins[k]` is the same data as a pinned int16 / float32 tensor).  Output buffers may be
    shared between stacks that are at least two apart (double buffering).
    schedule: optional list of chunk sizes (slices) in processing order, applied per stack.
    out_dtype: np.float32 (the reference's return type) or np.uint16 -- the 16-bit export
    ``uint16(clip(rint(x * 65535), 0, 65535))`` formed on the device (`mdimg_export_u16`), which halves the
    device-to-host bytes; metrics, validation rows and labels are those of the float32 image either way.
    trace: optional list that receives one record per chunk -- (worker, stack, chunk, slices, [5 CUDA events: copy-in
    start / end, compute start / end, copy-out end]) -- for timeline tools (tools/e2e_trace.py).
    Returns a list of (enhanced host array, StackResult without device pixels)."""
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.dtype(np.float32), np.dtype(np.uint16)):
        raise ValueError("out_dtype must be float32 or uint16")
    as_u16 = out_dtype == np.dtype(np.uint16)
    t_dtype = torch.int16 if as_u16 else torch.float32          # uint16 bit pattern in an int16 tensor
    ops = ops or get_ops()
    dev = ops.device
    nstk = len(raw_hosts)
    srcs, outs, packs, packs_dev, jobs, labels = [], [], [], [], [], []
    for k, raw_host in enumerate(raw_hosts):
        n, h, w = raw_host.shape
        pin = pinned_ins[k] if pinned_ins is not None else None
        if pin is not None:
            src_t = pin
        elif raw_host.dtype == np.uint16:
            src_t = torch.from_numpy(raw_host.view(np.int16))
        else:
            src_t = torch.from_numpy(np.ascontiguousarray(raw_host, np.float32))
        out_np = out_hosts[k] if out_hosts is not None else None
        pout = pinned_outs[k] if pinned_outs is not None else None
        if out_np is not None:
            if out_np.dtype != out_dtype or out_np.shape != (n, h, w):
                raise ValueError(f"output array {k}: expected {out_dtype} {(n, h, w)}, got {out_np.dtype} {out_np.shape}")
            out_t = torch.from_numpy(out_np.view(np.int16) if as_u16 else out_np)
        else:
            out_t = pout if pout is not None else torch.empty((n, h, w), dtype=t_dtype, pin_memory=True)
            if out_t.dtype != t_dtype:
                raise ValueError(f"pinned output {k}: expected a {t_dtype} tensor")
        srcs.append(src_t)
        outs.append(out_t)
        packs.append(torch.empty((n, PACK_COLS), dtype=torch.float64, pin_memory=True))
        packs_dev.append(torch.empty((n, PACK_COLS), dtype=torch.float64, device=dev))
        spans = _spans_of(n, chunk or default_chunk(h, w), schedule)
        labels.append([None] * len(spans))
        jobs.extend((k, j, a, b) for j, (a, b) in enumerate(spans))
    workers = max(1, min(workers, len(jobs)))
    caller = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(caller)
    errors: List[BaseException] = []
    next_job = [0]
    queue_lock = threading.Lock()

    def worker(wk: int) -> None:
        try:
            with torch.cuda.device(dev):
                main, copy_in, copy_out = _host_streams(dev, wk)

                def take():                       # chunks are handed out dynamically, in order
                    with queue_lock:
                        i = next_job[0]
                        next_job[0] += 1
                    return jobs[i] if i < len(jobs) else None

                copy_in.wait_event(ready)
                main.wait_event(ready)
                job = take()
                with torch.cuda.stream(main):
                    while job is not None:
                        # a worker claims its next chunk only when it is done with the current one
                        # (the safeguard read-backs pace the host thread with the GPU), so chunks
                        # are balanced dynamically; its copy-in overlaps the other workers' compute
                        k, j, a, b = job
                        tr = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if trace is not None else None
                        with torch.cuda.stream(copy_in):
                            if tr:
                                tr[0].record(copy_in)
                            raw_d = srcs[k][a:b].to(dev, non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(copy_in)
                            if tr:
                                tr[1].record(copy_in)
                        main.wait_event(ev)
                        raw_d.record_stream(main)
                        if tr:
                            tr[2].record(main)
                        enh, packed, lab = process_chunk(ops, raw_d, plan, True)
                        if as_u16:
                            enh = ops.export_u16(enh)
                        packs_dev[k][a:b].copy_(packed)          # device-resident copy of the rows (gather send buffer)
                        done = torch.cuda.Event(enable_timing=tr is not None)
                        done.record(main)
                        with torch.cuda.stream(copy_out):
                            copy_out.wait_event(done)
                            outs[k][a:b].copy_(enh, non_blocking=True)
                            packs[k][a:b].copy_(packed, non_blocking=True)
                            enh.record_stream(copy_out)
                            packed.record_stream(copy_out)
                            if tr:
                                tr[3] = done
                                tr[4].record(copy_out)
                                trace.append((wk, k, j, b - a, tr))
                        labels[k][j] = lab
                        job = take()
                copy_out.synchronize()
                main.synchronize()
        except BaseException as exc:  # noqa: BLE001 - re-raised on the caller's thread
            errors.append(exc)

    if workers == 1:
        worker(0)
    else:
        threads = [threading.Thread(target=worker, args=(wk,), daemon=True) for wk in range(workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    results = []
    for k in range(nstk):
        flat = [lab for part in labels[k] for lab in (part or [])]
        out_np = outs[k].numpy()
        results.append((out_np.view(np.uint16) if as_u16 else out_np,
                        StackResult(enhanced=None, packed=packs[k].numpy().copy(), labels=flat,
                                    packed_dev=packs_dev[k])))
    return results


def process_stack_host(raw_host: np.ndarray, plan, chunk: Optional[int] = None,
                       out_host: Optional[np.ndarray] = None, ops: Optional[StackOps] = None,
                       pinned_in: Optional[torch.Tensor] = None, pinned_out: Optional[torch.Tensor] = None,
                       workers: int = 2, schedule: Optional[List[int]] = None, out_dtype=np.float32):
    """One stack through `process_stacks_host`.  Returns (enhanced host array, StackResult)."""
    return process_stacks_host([raw_host], plan, chunk=chunk, out_hosts=[out_host], ops=ops,
                               pinned_ins=[pinned_in], pinned_outs=[pinned_out], workers=workers,
                               schedule=schedule, out_dtype=out_dtype)[0]


_host_stream_cache: dict = {}


def _host_streams(device: torch.device, k: int):
    """(compute, copy-in, copy-out) streams of host worker k.  The compute streams get descending
    priorities: with equal priorities the workers' chunks advance in lock step and finish in waves,
    and the copy-out of the whole last wave (one chunk per worker) overlaps nothing; with
    priorities the chunks complete one after another and only the last chunk's copy-out is exposed."""
    key = (device.index, k)
    if key not in _host_stream_cache:
        lo, hi = 0, -5
        try:
            lo, hi = torch.cuda.Stream.priority_range()        # (least, greatest), e.g. (0, -5)
        except Exception:  # noqa: BLE001
            pass
        prio = max(hi, lo - k) if hi < lo else lo
        _host_stream_cache[key] = (torch.cuda.Stream(device, priority=prio), torch.cuda.Stream(device),
                                   torch.cuda.Stream(device, priority=hi))
    return _host_stream_cache[key]


def save_stack_visuals(original: torch.Tensor, enhanced: torch.Tensor, out_dir: str, base_name: str,
                       workers: int = 8, gap: int = 8) -> List[str]:
    """Report mosaics for a whole stack (the reference writes one matplotlib figure
    per image, pipeline/dicom_io.py:99-126): one launch quantises every before | after pair on the
    GPU, the PNG encoding (zlib) runs on a host thread pool off the GPU's critical path.
    Returns the file paths ``<out_dir>/<base_name>_<index>_before_after.png``."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from . import png
    ops = get_ops(original.device)
    mosaics = ops.mosaic(original, enhanced, gap=gap).cpu().numpy()
    os.makedirs(out_dir, exist_ok=True)
    paths = [os.path.join(out_dir, f"{base_name}_{i:05d}_before_after.png") for i in range(mosaics.shape[0])]
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:      # zlib releases the GIL
        list(pool.map(lambda t: png.write_gray8(t[0], t[1]), zip(paths, mosaics)))
    return paths


def score_plans(images: torch.Tensor, plans, ops: Optional[StackOps] = None):
    """Tuning-loop batch (reference: pipeline/tools.py:95-183 called once per
    candidate and image by the LLM tuner): evaluate K candidate plans on N normalised images.
    The metrics of the originals are computed once and shared by every candidate; each candidate is
    one stack-wide enhancement + metrics + SSIM/PSNR pass.  Returns (scores [K, N] float64,
    results: list of K StackResult without pixels)."""
    ops = ops or get_ops(images.device)
    if images.dtype != torch.float32:
        raise ValueError("score_plans expects a float32 [N, H, W] stack normalised to [0, 1]")
    eng = Engine(ops)
    n = images.shape[0]
    rows_b = ops.metrics(images, with_niqe=True)
    scores = np.zeros((len(plans), n), np.float64)
    results = []
    for k, plan in enumerate(plans):
        res = eng.enhance_from_params(images, plan, rows_before=rows_b, on_error="flag")
        fr = ops.fullref(images, res.image)
        packed = torch.zeros((n, PACK_COLS), dtype=torch.float64, device=ops.device)
        packed[:, :ROW_COLS] = rows_b
        packed[:, ROW_COLS:2 * ROW_COLS] = res.rows_after
        packed[:, 2 * ROW_COLS:2 * ROW_COLS + 2] = fr
        host = packed.cpu().numpy()
        host[:, 2 * ROW_COLS + 2] = res.halo
        host[:, 2 * ROW_COLS + 3] = res.noise_guard
        host[:, 2 * ROW_COLS + 4] = res.over_processed
        if res.tv_iterations is not None:
            host[:, 2 * ROW_COLS + 5] = res.tv_iterations
        for i in res.errors:
            host[i, 2 * ROW_COLS + 6] = 1.0
        sr = StackResult(enhanced=None, packed=host, labels=res.labels)
        for i in range(n):
            # tool_score_plan returns -100 when validation failed with an error (tools.py:173-174)
            scores[k, i] = -100.0 if sr.failed[i] else sr.score(i)[0]
        results.append(sr)
    return scores, results
