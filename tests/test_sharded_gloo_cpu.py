"""world_size-2/3 gloo tests of the multi-GPU host logic (row partitioning, the packed all-gather
layout, rank-major merge order) on CPU.  The local index and the merge are the ORACLE here
(injected, test-only); on GPUs the same class runs the CUDA index and the K4 kernel."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import oracle  # noqa: E402
from c99_vectordb_b200.sharded import ShardedIndexFlat, shard_range, _packed_layout  # noqa: E402


class OracleLocalIndex:
    """add_with_ids / search on the CPU oracle (what one rank's CUDA index does)."""

    def __init__(self, d, metric):
        self.d, self.metric = d, metric
        self.rows = np.zeros((0, d), np.float32)
        self.ids = np.zeros((0,), np.int64)

    def add_with_ids(self, x, ids):
        self.rows = np.concatenate([self.rows, x])
        self.ids = np.concatenate([self.ids, ids])

    def search(self, q, k):
        return oracle.search(self.metric, self.rows, q, k, ids=self.ids)


class ApproxLocalIndex(OracleLocalIndex):
    """The row-sharded batch protocol of b200_index_search_shard_dev restated on the CPU: an approximate pass
    (exact score + bounded noise, standing in for the bf16 GEMM) emits the rows above a threshold that aims at
    c*k/world candidates, the candidates are re-ranked exactly, and the shard reports the bound no other row of it can
    beat.  Small candidate budgets make some certificates fail, which exercises the widened retry and the exact
    fallback."""

    EPS = 0.02

    def __init__(self, d, metric):
        super().__init__(d, metric)
        self.calls = []

    def search_shard(self, q, k, world, widen):
        self.calls.append((q.shape[0], widen))
        nq, n = q.shape[0], self.rows.shape[0]
        big = np.finfo(np.float32).max
        D = np.full((nq, k), -big if self.metric == 0 else big, np.float32)
        I = np.full((nq, k), -1, np.int64)
        B = np.full(nq, -np.inf if self.metric == 0 else np.inf, np.float32)
        if n == 0:
            return D, I, B
        if widen >= 2:  # the exact scan: this shard's true top k, a bound that excludes nothing
            D, I = oracle.search(self.metric, self.rows, q, k, ids=self.ids)
            return D, I, B
        rng = np.random.default_rng(1234 + n + widen)
        for i in range(nq):
            exact = (self.rows @ q[i]) if self.metric == 0 else ((self.rows - q[i]) ** 2).sum(axis=1)
            approx = exact + rng.uniform(-self.EPS, self.EPS, size=n).astype(np.float32)
            want = min(n, max(2, int((6 if widen else 1.2) * k / world)))
            order = np.sort(approx)
            if self.metric == 0:
                theta = order[n - want] - 1e-6 if want < n else -np.inf
                cand = np.nonzero(approx > theta)[0]
                B[i] = theta + self.EPS
            else:
                theta = order[want - 1] + 1e-6 if want < n else np.inf
                cand = np.nonzero(approx < theta)[0]
                B[i] = theta - self.EPS
            Dc, Ic = oracle.search(self.metric, self.rows[cand], q[i:i + 1], k, ids=self.ids[cand])
            D[i], I[i] = Dc[0], Ic[0]
        if widen == 1:  # an overflowed candidate list proves nothing: this query must reach the exact stage
            B[0] = np.inf if self.metric == 0 else -np.inf
        return D, I, B


def _batch_worker(rank, world, port, metric, n, d, k, nq, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = oracle.synth_rows(n // 2, d, 9)
        db = np.concatenate([base, base, oracle.synth_rows(n - 2 * (n // 2), d, 10)])  # cross-shard exact ties
        ids = np.arange(n, dtype=np.int64) * 7 + 3
        q = oracle.synth_rows(nq, d, 8)
        local = ApproxLocalIndex(d, metric)
        idx = ShardedIndexFlat(d, metric, local_index=local, merge_fn=oracle.merge_topk, device="cpu")
        idx.add_with_ids(db, ids)
        D, I = idx.search(q, k)
        Dw, Iw = oracle.search(metric, db, q, k, ids=ids)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
        np.save(os.path.join(out_dir, f"calls_{rank}.npy"), np.asarray(local.calls, dtype=np.int64))
        np.save(os.path.join(out_dir, f"unc_{rank}.npy"), np.asarray([idx.last_batch_uncertified]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,metric,n,k,nq", [(2, 0, 4001, 10, 24), (3, 1, 3000, 5, 16), (2, 1, 7, 5, 4)])
def test_sharded_batch_certificate_after_merge(tmp_path, world, metric, n, k, nq):
    """Batches: approximate candidates per shard, certificate over all shards after the merge, widened retry, exact
    fallback — the answer equals the unsharded oracle whatever the approximate pass did."""
    mp.spawn(_batch_worker, args=(world, _free_port(), metric, n, 24, k, nq, str(tmp_path)), nprocs=world, join=True)
    calls = [np.load(tmp_path / f"calls_{r}.npy") for r in range(world)]
    for c in calls[1:]:
        np.testing.assert_array_equal(c, calls[0])  # every rank took the same decisions
    assert calls[0][0].tolist() == [nq, 0]
    unc = int(np.load(tmp_path / "unc_0.npy")[0])
    if n > 1000:
        assert 0 < unc, "the small candidate budget was meant to leave some queries uncertified"
        assert len(calls[0]) >= 2 and calls[0][1].tolist() == [unc, 1]
        assert len(calls[0]) == 3 and calls[0][2].tolist() == [1, 2], "the query whose widened list 'overflowed' goes to the exact scan"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, metric, n, d, k, nq, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = oracle.synth_rows(n // 2, d, 9)
        db = np.concatenate([base, base, oracle.synth_rows(n - 2 * (n // 2), d, 10)])  # cross-shard exact ties
        ids = np.arange(n, dtype=np.int64) * 7 + 3
        q = oracle.synth_rows(nq, d, 8)
        idx = ShardedIndexFlat(d, metric, local_index=OracleLocalIndex(d, metric), merge_fn=oracle.merge_topk, device="cpu")
        idx.add_with_ids(db, ids)
        lo, hi = shard_range(n, world, rank)
        assert idx.local.rows.shape[0] == hi - lo and idx.ntotal == n
        D, I = idx.search(q, k)
        Dw, Iw = oracle.search(metric, db, q, k, ids=ids)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
        np.save(os.path.join(out_dir, f"I_{rank}.npy"), I)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,metric,n,k,nq", [(2, 0, 1001, 10, 3), (2, 1, 1001, 10, 1), (3, 1, 50, 40, 2), (2, 0, 3, 5, 2)])
def test_sharded_search_equals_unsharded(tmp_path, world, metric, n, k, nq):
    mp.spawn(_worker, args=(world, _free_port(), metric, n, 24, k, nq, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"I_{r}.npy") for r in range(world)]
    for o in outs[1:]:
        np.testing.assert_array_equal(o, outs[0])  # every rank holds the merged answer


def _query_exchange_worker(rank, world, port, metric, n, d, k, nq, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = oracle.synth_rows(n, d, 9)
        ids = np.arange(n, dtype=np.int64)
        q = oracle.synth_rows(nq, d, 8)
        idx = ShardedIndexFlat(d, metric, local_index=OracleLocalIndex(d, metric), merge_fn=oracle.merge_topk, device="cpu")
        idx.query_allgather_min_bytes = 0  # every host batch: upload 1/world of the queries, all-gather the rest
        idx.add_with_ids(db, ids)
        # this rank may only look at ITS slice of the host queries: everything else is poisoned here and must arrive
        # from the rank that owns it
        up = -(-nq // world)
        mine = q.copy()
        mine[: rank * up] = np.nan
        mine[(rank + 1) * up:] = np.nan
        D, I = idx.search(mine, k)
        Dw, Iw = oracle.search(metric, db, q, k, ids=ids)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 7), (3, 5), (2, 1), (3, 2)])
def test_host_batches_upload_one_slice_per_rank(tmp_path, world, nq):
    """Host batches: each rank stages and uploads rows [rank * ceil(nq / world), ...) of the queries only, one
    all-gather completes them everywhere (ragged and empty slices included)."""
    mp.spawn(_query_exchange_worker, args=(world, _free_port(), 0, 500, 24, 6, nq, str(tmp_path)), nprocs=world, join=True)


def test_shard_range_covers_rows_once():
    for n in (0, 1, 7, 8, 9, 100_000_000):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi >= lo for lo, hi in spans)


def test_packed_layout_alignment():
    for nq, k in ((1, 10), (3, 7), (10000, 100), (1, 1)):
        total, off_d, ib = _packed_layout(nq, k)
        assert off_d == nq * k * 8 and off_d % 8 == 0 and total % 16 == 0 and total >= nq * k * 12
