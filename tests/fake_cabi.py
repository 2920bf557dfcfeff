"""A stand-in for the index entry points of the C ABI, for CPU tests of the HOST logic in
c99_vectordb_b200/index.py and memo_adapter.py (file headers, id maps, argument marshalling, error
conversion).  TEST INFRASTRUCTURE ONLY — the product has no CPU path; this object is injected into
`_cabi._lib` by tests/test_index_host_cpu.py and never ships.  Host-only entry points of the real
library (b200_hash_embed, b200_py_hash_seed0, b200_last_error …) are forwarded to it.

Arithmetic: brute force in float64, rounded to float32 — good enough for ranking checks modulo near
ties; bit-exact parity is what the GPU tests are for."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


def _arr(ptr, n, ctype, dtype):
    if n == 0:
        return np.zeros((0,), dtype=dtype)
    addr = ptr if isinstance(ptr, int) else ptr.value
    return np.ctypeslib.as_array((ctype * n).from_address(addr)).view(dtype)


class _Idx:
    def __init__(self, d, metric, store):
        self.d, self.metric, self.store = d, metric, store
        self.rows = np.zeros((0, d), np.float32)
        self.ids = None
        self.opts = {}


class FakeLib:
    def __init__(self, real):
        self._real = real
        self._tab: dict[int, _Idx] = {}
        self._next = 1
        self._err = b""

    def __getattr__(self, name):  # host-only entry points and anything not faked
        return getattr(self._real, name)

    # ---- helpers ----
    def _fail(self, msg):
        self._err = msg.encode()
        return 1

    def b200_last_error(self):
        if self._err:  # the last failure came from a faked entry point
            e, self._err = self._err, b""
            return e
        return self._real.b200_last_error()

    def _ix(self, h) -> _Idx:
        return self._tab[h.value if hasattr(h, "value") else int(h)]

    # ---- lifetime ----
    def b200_index_create(self, out_ref, d, metric, store, device):
        if d <= 0:
            return self._fail(f"d must be positive, got {d}")
        out_ref._obj.value = self._next
        self._tab[self._next] = _Idx(int(d), int(metric), int(store))
        self._next += 1
        return 0

    def b200_index_destroy(self, h):
        self._tab.pop(h.value, None)
        return 0

    def b200_index_reset(self, h):
        ix = self._ix(h)
        ix.rows, ix.ids = np.zeros((0, ix.d), np.float32), None
        return 0

    def b200_index_reserve(self, h, n):
        return 0

    def b200_index_sync(self, h):
        return 0

    def b200_index_set_option(self, h, name, value):
        self._ix(h).opts[name.decode()] = int(value)
        return 0

    def b200_index_get_option(self, h, name, out_ref):
        out_ref._obj.value = self._ix(h).opts.get(name.decode(), 0)
        return 0

    def b200_index_ntotal(self, h):
        return self._ix(h).rows.shape[0]

    def b200_index_launch_count(self, h):
        return 0

    # ---- add ----
    def _append(self, ix, x, ids, normalize):
        x = np.array(x, dtype=np.float32)
        if normalize:
            nrm = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
            x = np.where(nrm <= 1e-8, 0, x / np.maximum(nrm, 1e-30)).astype(np.float32)
        if (ids is None) != (ix.ids is None) and ix.rows.shape[0]:
            return self._fail("cannot mix add() and add_with_ids() on one index")
        ix.rows = np.concatenate([ix.rows, x])
        if ids is not None:
            ix.ids = np.concatenate([ix.ids if ix.ids is not None else np.zeros((0,), np.int64), np.array(ids, np.int64)])
        return 0

    def b200_index_add(self, h, x, n, ids, normalize):
        ix = self._ix(h)
        X = _arr(x, n * ix.d, C.c_float, np.float32).reshape(n, ix.d)
        I = _arr(ids, n, C.c_int64, np.int64) if ids else None
        return self._append(ix, X, I, normalize)

    def b200_index_add_texts(self, h, blob, offsets, n, ids, first_id, skip_blank, normalize, with_ids, n_added_ref):
        """Host restatement of K6 for the CPU tests of pack_texts / add_texts: ASCII lower-casing, blank = only ASCII
        white space, buckets from the library's own host embedder (b200_hash_embed)."""
        ix = self._ix(h)
        off = _arr(offsets, n + 1, C.c_int64, np.int64)
        data = bytes(blob)
        recs = [data[off[i]:off[i + 1]] for i in range(n)]
        space = set(b"\t\n\x0b\x0c\r\x1c\x1d\x1e\x1f ")
        keep = [i for i, r in enumerate(recs) if not (skip_blank and all(c in space for c in r))]
        low = [bytes(c + 32 if 65 <= c <= 90 else c for c in recs[i]) for i in keep]
        o2 = np.zeros(len(keep) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in low], out=o2[1:])
        out = np.zeros((len(keep), ix.d), dtype=np.float32)
        if keep:
            rc = self._real.b200_hash_embed(b"".join(low), o2.ctypes.data, len(keep), ix.d, out.ctypes.data)
            if rc:
                return self._fail("b200_hash_embed failed")
        I = None
        if ids:
            I = _arr(ids, n, C.c_int64, np.int64)[keep]
        elif with_ids:
            I = np.asarray(keep, dtype=np.int64) + first_id
        n_added_ref._obj.value = len(keep)
        if not keep:
            return 0
        return self._append(ix, out, I, normalize)

    def b200_index_add_file(self, h, path, rows_off, n, ids_off, normalize):
        ix = self._ix(h)
        size = os.path.getsize(path)
        if rows_off + n * ix.d * 4 > size or (ids_off >= 0 and ids_off + n * 8 > size):
            return self._fail(f"read error: {path.decode()} is shorter than its header promises")
        with open(path, "rb") as f:
            f.seek(rows_off)
            X = np.frombuffer(f.read(n * ix.d * 4), dtype="<f4").reshape(n, ix.d)
            I = None
            if ids_off >= 0:
                f.seek(ids_off)
                I = np.frombuffer(f.read(n * 8), dtype="<i8")
        return self._append(ix, X, I, normalize)

    def b200_index_write_file(self, h, path, rows_off, ids_off):
        ix = self._ix(h)
        with open(path, "r+b") as f:
            f.seek(rows_off)
            f.write(ix.rows.astype("<f4").tobytes())
            if ids_off >= 0:
                f.seek(ids_off)
                f.write(self._ids(ix).astype("<i8").tobytes())
        return 0

    # ---- read back ----
    @staticmethod
    def _ids(ix):
        return ix.ids if ix.ids is not None else np.arange(ix.rows.shape[0], dtype=np.int64)

    def b200_index_get_ids(self, h, out):
        ix = self._ix(h)
        n = ix.rows.shape[0]
        _arr(out, n, C.c_int64, np.int64)[:] = self._ids(ix)
        return 0

    def b200_index_get_rows(self, h, row0, n, out):
        ix = self._ix(h)
        if row0 < 0 or n < 0 or row0 + n > ix.rows.shape[0]:
            return self._fail(f"row range [{row0},+{n}) out of bounds")
        _arr(out, n * ix.d, C.c_float, np.float32)[:] = ix.rows[row0:row0 + n].reshape(-1)
        return 0

    # ---- search ----
    def _search(self, ix, q, nq, k, D, I, allow_rows):
        if k <= 0:
            return self._fail(f"k must be positive, got {k}")
        Q = _arr(q, nq * ix.d, C.c_float, np.float32).reshape(nq, ix.d).astype(np.float64)
        if ix.opts.get("normalize_queries"):
            Q = Q / np.maximum(np.linalg.norm(Q, axis=1, keepdims=True), 1e-30)
        Dv = _arr(D, nq * k, C.c_float, np.float32).reshape(nq, k)
        Iv = _arr(I, nq * k, C.c_int64, np.int64).reshape(nq, k)
        rows = ix.rows.astype(np.float64)
        ids = self._ids(ix)
        for i in range(nq):
            if ix.metric == 0:
                s = rows @ Q[i]
                order = np.argsort(-s, kind="stable")
            else:
                s = ((rows - Q[i]) ** 2).sum(axis=1)
                order = np.argsort(s, kind="stable")
            if allow_rows is not None:
                order = order[allow_rows[order]]
            m = min(k, order.shape[0])
            Dv[i, :m] = s[order[:m]]
            Iv[i, :m] = ids[order[:m]]
            Dv[i, m:] = -np.finfo(np.float32).max if ix.metric == 0 else np.finfo(np.float32).max
            Iv[i, m:] = -1
        return 0

    def b200_index_search(self, h, q, nq, k, D, I):
        return self._search(self._ix(h), q, nq, k, D, I, None)

    def b200_index_search_masked(self, h, q, nq, k, mask, D, I):
        ix = self._ix(h)
        n = ix.rows.shape[0]
        words = _arr(mask, (n + 31) // 32, C.c_uint32, np.uint32)
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)
        return self._search(ix, q, nq, k, D, I, bits)

    def b200_index_search_ids_allowed(self, h, q, nq, k, allowed, m, D, I):
        ix = self._ix(h)
        a = _arr(allowed, m, C.c_int64, np.int64) if m else np.zeros((0,), np.int64)
        return self._search(ix, q, nq, k, D, I, np.isin(self._ids(ix), a))

    # ---- stand-alone stages ----
    def b200_normalize_rows(self, x, n, d, device):
        X = _arr(x, n * d, C.c_float, np.float32).reshape(n, d)
        for r in range(n):  # memo_cli.py:131-135
            nrm = float(np.linalg.norm(X[r]))
            X[r] = 0 if nrm <= 1e-8 else X[r] / np.float32(nrm)
        return 0
