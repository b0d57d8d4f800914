"""`shard.process_cohort` on the CPU (gloo, world_size 2): sharding of a cohort of volumes into
contiguous per-rank slice ranges, the row gather in cohort order (equal and unequal shards, whole
cohort on every rank and pre-sharded inputs), label gather and persistence on rank 0.  The CUDA
pipeline is replaced by a stand-in `processor` (the GPU leg is tests/test_gpu_dropin.py)."""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from mdimg_b200.shard import cohort_spans, slice_range


def test_cohort_spans_partition_the_cohort():
    for sizes in ([2048] * 4, [5, 0, 3, 9], [1], [7, 7, 7], [0, 0]):
        total = sum(sizes)
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                spans = cohort_spans(sizes, r, world)
                cnt = sum(b - a for _, a, b in spans)
                a, b = slice_range(total, r, world)
                assert cnt == b - a
                for v, lo, hi in spans:
                    assert 0 <= lo < hi <= sizes[v]
                    start = sum(sizes[:v])
                    seen.extend(range(start + lo, start + hi))
            assert seen == list(range(total))           # cohort order, every slice exactly once
    # C4: 4 volumes x 2048 slices over 8 ranks -> 1024 slices per rank, half a volume each
    assert cohort_spans([2048] * 4, 3, 8) == [(1, 1024, 2048)]


def _fake_processor(stack, plan):
    """Stand-in for the CUDA pipeline: row i of a slice = its mean, its first pixel, the plan tag."""
    from mdimg_b200.batch import PACK_COLS, StackResult
    n = stack.shape[0]
    packed = np.zeros((n, PACK_COLS))
    packed[:, 0] = stack.reshape(n, -1).mean(axis=1)
    packed[:, 1] = stack[:, 0, 0]
    packed[:, 2] = plan
    labels = [[f"v{int(stack[i, 0, 0])}"] for i in range(n)]
    return stack * 2, StackResult(enhanced=None, packed=packed, labels=labels, packed_dev=torch.from_numpy(packed))


def _cohort(sizes):
    vols, k = [], 0
    for n in sizes:
        v = np.zeros((n, 4, 4), np.float32)
        for i in range(n):
            v[i] = k
            k += 1
        vols.append(v)
    return vols


def _worker(rank, world, sizes, sharded, port, db, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MDIMG_DB_PATH"] = db
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mdimg_b200.shard import process_cohort
    vols = _cohort(sizes)
    if sharded:                                        # every rank holds only its own volumes
        vols = [v for i, v in enumerate(vols) if i % world == rank]
    res = process_cohort(vols, 7.0, already_sharded=sharded, processor=_fake_processor, gather_label_lists=True)
    q.put((rank, (res.rows.numpy().copy(), res.counts, res.labels,
                  [(v, a, b, float(enh[0, 0, 0]) if len(enh) else None) for v, a, b, enh, _ in res.local])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sizes,sharded", [([4, 4], False), ([5, 0, 4], False), ([3, 6], True), ([2, 2, 2, 2], True)])
def test_process_cohort_gloo_world2(sizes, sharded, tmp_path):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 1500) + 7 * len(sizes) + sum(sizes) + (50 if sharded else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, sizes, sharded, port, str(tmp_path / "runs.db"), q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(sizes)
    if sharded:     # rank r holds volumes r, r + 2, ...: cohort order = rank 0's slices, then rank 1's
        ids = [np.concatenate([np.arange(sum(sizes[:i]), sum(sizes[:i + 1])) for i in range(len(sizes)) if i % 2 == r] or
                              [np.zeros(0)]) for r in range(2)]
        order = np.concatenate(ids)
        counts = [len(ids[0]), len(ids[1])]
    else:
        order = np.arange(total)
        counts = [b - a for a, b in (slice_range(total, r, 2) for r in range(2))]
    for r in range(2):
        rows, got_counts, labels, local = results[r]
        assert got_counts == counts
        assert rows.shape[0] == total
        np.testing.assert_array_equal(rows[:, 0], order)            # every rank holds the whole cohort's rows, in order
        np.testing.assert_array_equal(rows[:, 1], order)
        assert np.all(rows[:, 2] == 7.0)
        assert labels == [[f"v{int(i)}"] for i in order]
        assert sum(b - a for _, a, b, _ in local) == counts[r]     # pixels stay with the rank that produced them
    np.testing.assert_array_equal(results[0][0], results[1][0])
