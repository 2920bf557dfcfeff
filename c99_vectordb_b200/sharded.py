"""Row-sharded flat index across the GPUs of one box: one process per GPU (torch.distributed).

The reference has no distributed code (single process, memo_cli.py:883-949); this is the multi-GPU
restatement north_star specifies for index.search (memo_cli.py:292): rows are split into contiguous
ranges, rank g owning [g*ceil(N/G), (g+1)*ceil(N/G)); every rank runs the scan kernel over its
shard (ids are translated to GLOBAL record ids inside the kernel), the per-rank best-first lists
are exchanged either inside the scan kernel itself (fused: peer-memory stores over NVLink + flags +
in-kernel merge, enable_fused_exchange) or with ONE NCCL all-gather of a packed (I,D) buffer followed
by the K4 merge kernel on every rank.  Contiguous ranges make "lower rank first, then earlier list
position" equal to the global tie rule (smaller row position first).

torch is plumbing only: device buffers, the stream, and the process group.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from . import index as _ix


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; the last ranks may be short or empty."""
    per = -(-int(n_total) // int(world)) if n_total > 0 else 0
    lo = min(n_total, rank * per)
    hi = min(n_total, (rank + 1) * per)
    return lo, hi


def _packed_layout(nq: int, k: int) -> tuple[int, int, int]:
    """Per-rank exchange buffer: int64 I[nq,k] first (8-byte aligned), then float32 D[nq,k];
    padded to 16 bytes.  Returns (bytes, offset_of_D, bytes_of_I)."""
    ib = nq * k * 8
    db = nq * k * 4
    total = (ib + db + 15) // 16 * 16
    return total, ib, ib


class ShardedIndexFlat:
    """faiss-shaped add/search over a row-sharded database.

    `local_index` / `merge_fn` / `device` exist so the host logic (partitioning, exchange layout,
    rank-major merge order) can be exercised on CPU with the gloo backend in tests; the defaults
    are the CUDA index and the K4 kernel and nothing else ships.
    """

    def __init__(self, d: int, metric: int = _ix.METRIC_L2, *, store: str = "f32", normalize: bool = False,
                 group=None, local_index=None, merge_fn=None, device=None):
        import torch
        import torch.distributed as dist

        self.d, self.metric_type = int(d), int(metric)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if local_index is None:
            dev_index = torch.cuda.current_device()
            self.device = torch.device("cuda", dev_index)
            local_index = _ix.IndexIDMap2(_ix.IndexFlat(d, metric, store=store, normalize=normalize, device=dev_index))
        else:
            self.device = torch.device(device or "cpu")
        self.local = local_index
        self._merge_fn = merge_fn
        self.ntotal_global = 0
        self.merge_launches = 0
        self._bufs = {}
        self._fused = False
        self._host_stage = {}
        self.last_batch_uncertified = None
        self.last_batch_stats = None
        # host batches of at least this many bytes: every rank uploads 1/world of the queries and the rest arrives by
        # an all-gather over NVLink (all ranks hold the same host queries; the PCIe links are the narrow part)
        self.query_allgather_min_bytes = 1 << 20

    # ---- add --------------------------------------------------------------------------------
    def add_with_ids(self, x: np.ndarray, ids: np.ndarray) -> None:
        """Every rank is handed the same [n,d] block; it keeps the rows of its contiguous range."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        assert x.ndim == 2 and x.shape[1] == self.d and ids.shape == (x.shape[0],)
        if self.ntotal_global != 0:
            raise RuntimeError("ShardedIndexFlat takes its rows in one add (contiguous ranges per rank)")
        lo, hi = shard_range(x.shape[0], self.world, self.rank)
        if hi > lo:
            self.local.add_with_ids(x[lo:hi], ids[lo:hi])
        self.ntotal_global = x.shape[0]

    def add(self, x: np.ndarray) -> None:
        self.add_with_ids(x, np.arange(np.shape(x)[0], dtype=np.int64))

    def add_synthetic(self, n_total: int, seed: int) -> None:
        """Each rank generates its own range of the counter-based database on its device; ids are the
        global row positions."""
        if self.ntotal_global != 0:
            raise RuntimeError("ShardedIndexFlat takes its rows in one add")
        lo, hi = shard_range(n_total, self.world, self.rank)
        if hi > lo:
            self.local.index.add_synthetic(hi - lo, seed, first_row=lo, with_ids=True, first_id=lo)
        self.ntotal_global = int(n_total)

    @property
    def ntotal(self) -> int:
        return self.ntotal_global

    @property
    def launch_count(self) -> int:
        base = getattr(self.local, "index", self.local)
        return int(getattr(base, "launch_count", 0)) + self.merge_launches

    # ---- fused exchange -----------------------------------------------------------------------
    def enable_fused_exchange(self) -> bool:
        """Switch the top-k exchange from NCCL all-gather + K4 to the fused form: the scan kernel's
        last CTA stores its local result into every rank's exchange buffer over NVLink (CUDA IPC peer
        mappings), flags it, waits for the peers and merges — one kernel per GPU per search.
        Collective: every rank must call it.  Returns False (and keeps NCCL) on a single rank."""
        import torch.distributed as dist

        if self.world == 1 or self.device.type != "cuda":
            return False
        import torch

        L = _cabi.load()
        ok, err = 1, ""
        peers = (C.c_void_p * self.world)()
        try:
            nbytes = 2 * self.world * int(L.b200_exchange_slot_bytes())
            handle = C.create_string_buffer(64)
            mine = C.c_void_p()
            _cabi.check(L.b200_ipc_alloc(C.byref(mine), nbytes, handle))
            raw = handle.raw
        except RuntimeError as e:  # keep the collective below balanced even if this rank failed
            ok, err, raw = 0, str(e), b"\0" * 64
        handles = [None] * self.world
        dist.all_gather_object(handles, raw, group=self.group)
        if ok:
            try:
                for g in range(self.world):
                    if g == self.rank:
                        peers[g] = mine.value
                    else:
                        p = C.c_void_p()
                        _cabi.check(L.b200_ipc_open(handles[g], C.byref(p)))
                        peers[g] = p.value
                _cabi.check(L.b200_index_set_exchange(self.local.index._h, self.world, self.rank, peers))
            except RuntimeError as e:
                ok, err = 0, str(e)
        # all ranks or none: a single rank without peer mappings would deadlock the others
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self._fused = bool(flag.item())
        if not self._fused and err:
            import sys

            print(f"[b200] fused exchange unavailable on rank {self.rank} ({err}); using NCCL", file=sys.stderr)
        return self._fused

    def _fused_ok(self, nq: int, k: int) -> bool:
        # The fused exchange lives in the scan kernel's tail and serves single queries (the latency
        # path).  Batches go through NCCL so that every shard can use the tensor-core path (K3), which
        # beats the scan from 2 queries on.  Every rank must hold rows (an empty shard launches no
        # kernel and nobody would flag for it).
        lo, hi = shard_range(self.ntotal_global, self.world, self.world - 1)
        return self._fused and nq == 1 and k <= 256 and hi > lo

    def check_exchange(self) -> None:
        """Raise if a fused exchange timed out (a peer GPU never delivered: the affected search returned padding
        only).  Call after synchronising the stream the searches ran on; search() does it for every call."""
        if self._fused and _cabi.load().b200_index_exchange_status(self.local.index._h) != 0:
            raise RuntimeError("fused exchange: a peer GPU did not deliver its results in time; the search returned no results")

    # ---- search -----------------------------------------------------------------------------
    def _buffers(self, nq: int, k: int):
        import torch

        key = (nq, k)
        if key not in self._bufs:
            nbytes, off_d, _ = _packed_layout(nq, k)
            mine = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            gathered = torch.zeros(self.world * nbytes, dtype=torch.uint8, device=self.device)
            D = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            self._bufs[key] = (mine, gathered, D, I, nbytes, off_d)
        return self._bufs[key]

    def search_device(self, q, k: int):
        """q: [nq,d] float32 tensor on this rank's device (same on every rank).  Returns (D, I)
        tensors holding the merged global result on every rank.  Enqueued on the current stream."""
        import torch
        import torch.distributed as dist

        nq, k = int(q.shape[0]), int(k)
        if self._fused_ok(nq, k):
            _, _, D, I, _, _ = self._buffers(nq, k)
            stream = torch.cuda.current_stream(self.device).cuda_stream or 1
            _cabi.check(_cabi.load().b200_index_search_exchange_dev(
                self.local.index._h, q.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), C.c_void_p(stream)))
            return D, I
        if nq >= 2 and k <= 256 and self.world > 1 and (self.device.type == "cuda" or hasattr(self.local, "search_shard")):
            return self._search_batch(q, k, widen=0)
        return self._search_gather(q, k)

    def _search_gather(self, q, k: int):
        """Exact local top k on every shard -> one all-gather -> K4 merge."""
        import torch
        import torch.distributed as dist

        nq = int(q.shape[0])
        mine, gathered, D, I, nbytes, off_d = self._buffers(nq, k)
        I_loc = mine[: nq * k * 8].view(torch.int64).view(nq, k)
        D_loc = mine[off_d: off_d + nq * k * 4].view(torch.float32).view(nq, k)
        self._local_search(q, k, D_loc, I_loc)
        if self.world == 1:
            return D_loc, I_loc
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        self._merge(gathered, nq, k, nbytes, off_d, D, I)
        return D, I

    # ---- batches: tensor-core path per shard, certificate after the merge -------------------------
    def _batch_buffers(self, nq: int, k: int):
        import torch

        key = ("batch", nq, k)
        if key not in self._bufs:
            ib, db, bb = nq * k * 8, nq * k * 4, nq * 4
            nbytes = (ib + db + bb + 15) // 16 * 16
            self._bufs[key] = (torch.zeros(nbytes, dtype=torch.uint8, device=self.device),
                               torch.zeros(self.world * nbytes, dtype=torch.uint8, device=self.device),
                               torch.empty((nq, k), dtype=torch.float32, device=self.device),
                               torch.empty((nq, k), dtype=torch.int64, device=self.device),
                               torch.zeros(nq, dtype=torch.int32, device=self.device),
                               torch.zeros(1, dtype=torch.int32, device=self.device), nbytes, ib, ib + db)
        return self._bufs[key]

    def _search_batch(self, q, k: int, widen: int):
        """Batched queries over row shards: every shard runs the tensor-core path with thresholds that
        aim at 1/world of the candidates (b200_index_search_shard_dev: bf16 GEMM + fused emit + exact re-rank of ITS
        rows only, nothing read back), the lists and per-query exclusion bounds travel in ONE all-gather, and the
        certificate is taken after the merge: a query is done when its merged k-th entry strictly beats every shard's
        bound.  The rest (identical on every rank, so no agreement step) is searched again: widen = 1 collects 3x the
        candidates, widen = 2 runs the exact scan on every shard (bounds that exclude nothing: always certified)."""
        import torch
        import torch.distributed as dist

        nq = int(q.shape[0])
        mine, gathered, D, I, unc, n_unc, nbytes, off_d, off_b = self._batch_buffers(nq, k)
        I_loc = mine[: nq * k * 8].view(torch.int64).view(nq, k)
        D_loc = mine[off_d: off_d + nq * k * 4].view(torch.float32).view(nq, k)
        B_loc = mine[off_b: off_b + nq * 4].view(torch.float32)
        cuda = self.device.type == "cuda"
        if cuda:
            L = _cabi.load()
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream or 1)
            _cabi.check(L.b200_index_search_shard_dev(self.local.index._h, q.data_ptr(), nq, k, self.world, int(widen),
                                                      D_loc.data_ptr(), I_loc.data_ptr(), B_loc.data_ptr(), stream))
        else:  # injected test index: numpy in, numpy out
            Dn, In, Bn = self.local.search_shard(q.numpy(), k, self.world, widen)
            D_loc.copy_(torch.from_numpy(np.ascontiguousarray(Dn)))
            I_loc.copy_(torch.from_numpy(np.ascontiguousarray(In)))
            B_loc.copy_(torch.from_numpy(np.ascontiguousarray(Bn, dtype=np.float32)))
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        if cuda:
            base = gathered.data_ptr()
            _cabi.check(L.b200_merge_certify_dev(
                self.metric_type, self.world, nq, k, self.ntotal_global, C.c_void_p(base + off_d), C.c_void_p(base),
                nbytes // 4, nbytes // 8, C.c_void_p(base + off_b), nbytes // 4, D.data_ptr(), I.data_ptr(),
                unc.data_ptr(), n_unc.data_ptr(), stream))
            self.merge_launches += 2
        else:
            g = gathered.view(self.world, nbytes)
            Ip = g[:, : nq * k * 8].contiguous().view(torch.int64).view(self.world, nq, k)
            Dp = g[:, off_d: off_d + nq * k * 4].contiguous().view(torch.float32).view(self.world, nq, k)
            Bp = g[:, off_b: off_b + nq * 4].contiguous().view(torch.float32).view(self.world, nq).numpy()
            Dm, Im = self._merge_fn(self.metric_type, Dp.numpy(), Ip.numpy())
            want = min(k, self.ntotal_global)
            if want > 0:
                kth = Dm[:, want - 1]
                if self.metric_type == _ix.METRIC_INNER_PRODUCT:
                    ok = (kth > -np.finfo(np.float32).max) & (kth[None, :] > Bp).all(axis=0)
                else:
                    ok = (kth < np.finfo(np.float32).max) & (kth[None, :] < Bp).all(axis=0)
            else:
                ok = np.ones(nq, dtype=bool)
            D.copy_(torch.from_numpy(Dm))
            I.copy_(torch.from_numpy(Im))
            unc.copy_(torch.from_numpy((~ok).astype(np.int32)))
            n_unc.fill_(int((~ok).sum()))
        bad_count = int(n_unc.item())  # one 4-byte read per batch: every rank sees the same count
        if widen == 0:
            self.last_batch_uncertified = bad_count  # of this batch's first attempt (bench.py reports it)
            if cuda:  # the pass brackets of the first attempt, before a retry overwrites them
                base = self.local.index
                self.last_batch_stats = {s: base.get_option("stat_gemm_" + s) for s in ("used", "pass1_us", "pass2_us", "rerank_us")}
        if bad_count and widen < 2:  # (after the exact scan there is nothing left to try: e.g. fewer valid rows than k)
            bad = torch.nonzero(unc, as_tuple=False).flatten()
            q_bad = q.index_select(0, bad).contiguous()
            # same protocol, widened (3x the candidates), then with the exact scan on every shard (bounds exclude nothing)
            D_bad, I_bad = self._search_batch(q_bad, k, widen=widen + 1)
            D, I = D.clone(), I.clone()  # the retry may have reused the cached buffers of this shape
            D.index_copy_(0, bad, D_bad)
            I.index_copy_(0, bad, I_bad)
        return D, I

    def search(self, x: np.ndarray, k: int):
        """Host query -> host result (every rank passes the same query and gets the same answer).
        Pinned staging buffers are cached per shape; one stream synchronisation per call."""
        import torch

        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        if self.device.type != "cuda":  # injected test index (gloo): same query exchange, no pinned staging
            if self.world > 1 and x.nbytes >= self.query_allgather_min_bytes:
                up = -(-x.shape[0] // self.world)
                q = self._upload_queries(x, torch.empty((up, self.d), dtype=torch.float32),
                                         torch.empty((up * self.world, self.d), dtype=torch.float32), True).contiguous()
            else:
                q = torch.from_numpy(x)
            D, I = self.search_device(q, k)
            return D.numpy().copy(), I.numpy().copy()
        if self._fused_ok(x.shape[0], int(k)):
            # latency path: ONE C call — pinned staging, H2D, the fused kernel (its last CTA writes the merged result
            # straight into pinned host memory), one synchronisation, the exchange status check.  It runs on the index's
            # own stream; the library orders that stream behind the one the handle used last.
            D = np.empty((x.shape[0], int(k)), dtype=np.float32)
            I = np.empty((x.shape[0], int(k)), dtype=np.int64)
            _cabi.check(_cabi.load().b200_index_search_exchange(self.local.index._h, x.ctypes.data, x.shape[0], int(k),
                                                                D.ctypes.data, I.ctypes.data))
            return D, I
        nq, k = x.shape[0], int(k)
        gather = self.world > 1 and x.nbytes >= self.query_allgather_min_bytes
        key = (nq, k, gather)
        st = self._host_stage.get(key)
        if st is None:
            up = -(-nq // self.world) if gather else nq  # query rows this rank uploads
            st = (torch.empty((up, self.d), dtype=torch.float32).pin_memory(),
                  torch.empty((up * self.world if gather else nq, self.d), dtype=torch.float32, device=self.device),
                  torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                  torch.empty((nq, k), dtype=torch.int64).pin_memory())
            self._host_stage[key] = st
        hq, dq, hD, hI = st
        dq = self._upload_queries(x, hq, dq, gather)
        D, I = self.search_device(dq, k)
        hD.copy_(D, non_blocking=True)
        hI.copy_(I, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.check_exchange()
        return hD.numpy().copy(), hI.numpy().copy()

    def _upload_queries(self, x: np.ndarray, hq, dq, gather: bool):
        """Host queries -> dq on this rank's device.  gather: this rank stages and uploads only rows
        [rank * up, (rank + 1) * up) (zero padded past nq) and one all-gather completes dq on every rank."""
        import torch.distributed as dist

        nq = x.shape[0]
        if not gather:
            hq.numpy()[...] = x
            dq.copy_(hq, non_blocking=True)
            return dq
        up = hq.shape[0]
        lo, hi = min(nq, self.rank * up), min(nq, (self.rank + 1) * up)
        h = hq.numpy()
        h[: hi - lo] = x[lo:hi]
        h[hi - lo:] = 0.0
        mine = dq[self.rank * up: (self.rank + 1) * up]
        mine.copy_(hq, non_blocking=True)
        dist.all_gather_into_tensor(dq, mine, group=self.group)
        return dq[:nq]

    # ---- pieces -----------------------------------------------------------------------------
    def _local_search(self, q, k, D_loc, I_loc) -> None:
        if self.device.type == "cuda":
            self.local.search_device(q, k, D=D_loc, I=I_loc)
        else:  # injected test index: numpy in, numpy out
            import torch

            Dn, In = self.local.search(q.numpy(), k)
            D_loc.copy_(torch.from_numpy(np.ascontiguousarray(Dn)))
            I_loc.copy_(torch.from_numpy(np.ascontiguousarray(In)))

    def _merge(self, gathered, nq, k, nbytes, off_d, D, I) -> None:
        import torch

        if self._merge_fn is not None:
            g = gathered.view(self.world, nbytes)
            Ip = g[:, : nq * k * 8].contiguous().view(torch.int64).view(self.world, nq, k)
            Dp = g[:, off_d: off_d + nq * k * 4].contiguous().view(torch.float32).view(self.world, nq, k)
            Dm, Im = self._merge_fn(self.metric_type, Dp.numpy(), Ip.numpy())
            D.copy_(torch.from_numpy(Dm))
            I.copy_(torch.from_numpy(Im))
            return
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        base = gathered.data_ptr()
        _cabi.check(_cabi.load().b200_merge_topk_dev(
            self.metric_type, self.world, nq, k, C.c_void_p(base + off_d), C.c_void_p(base), nbytes // 4, nbytes // 8,
            D.data_ptr(), I.data_ptr(), C.c_void_p(stream)))
        self.merge_launches += 1
