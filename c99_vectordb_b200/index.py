"""The faiss-shaped Python surface memo drives (memo_cli.py:13 `import faiss`), over the C ABI.

Exactly the symbols memo_cli.py touches plus the flat classes north_star names:

    IndexFlat / IndexFlatIP / IndexFlatL2      restated target of create_index(), memo_cli.py:244-248
    IndexHNSWFlat(d, M)                        constructor memo calls (:245); mapped to an exact flat
                                               L2 index (k = ntotal makes HNSW exhaustive anyway, :291)
    IndexIDMap / IndexIDMap2                   :248, isinstance checks :258
    .add_with_ids(x, ids) .add(x) .search(x,k) :282 :437 :292
    .ntotal .id_map  vector_to_array()         :266 :268 :289 :291 :473
    read_index / write_index                   :255 :361 :448
    normalize_L2                               faiss helper, the K1 kernel as a function

All arithmetic runs in the sm_100a kernels behind include/b200_flat.h.  There is no CPU path:
without the native library or a CUDA device every call raises RuntimeError.
Error convention follows faiss: RuntimeError for engine failures, AssertionError for shape checks.
"""
from __future__ import annotations

import ctypes as C
import os
from types import SimpleNamespace

import numpy as np

from . import _cabi

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def _default_device() -> int:
    for var in ("B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v is not None and v.strip().lstrip("-").isdigit():
            return int(v)
    return 0


class Int64Vector:
    """Stand-in for faiss's std::vector<int64> wrapper (`index.id_map`, memo_cli.py:268)."""

    def __init__(self, arr: np.ndarray):
        self._arr = np.ascontiguousarray(arr, dtype=np.int64)

    def size(self) -> int:
        return int(self._arr.shape[0])

    def at(self, i: int) -> int:
        return int(self._arr[i])

    def __len__(self) -> int:
        return self.size()


def vector_to_array(v) -> np.ndarray:
    """faiss.vector_to_array (memo_cli.py:268): a fresh numpy copy of the vector."""
    if isinstance(v, Int64Vector):
        return v._arr.copy()
    return np.array(v, copy=True)


def pack_row_mask(row_mask, ntotal: int) -> np.ndarray:
    """bool[ntotal] -> the uint32 bitmap include/b200_flat.h defines (bit r&31 of word r>>5)."""
    m = np.ascontiguousarray(row_mask, dtype=bool)
    assert m.shape == (ntotal,), "row_mask must have one entry per stored row"
    words = (ntotal + 31) // 32
    padded = np.zeros(words * 32, dtype=bool)
    padded[:ntotal] = m
    return np.packbits(padded, bitorder="little").view("<u4").copy()


def pack_texts(texts, ids=None, skip_blank: bool = True):
    """Records -> (utf-8 blob, int64 offsets[n+1], kept ids or None).  ASCII corpora travel as they are (one join, one
    encode; the device lower-cases and skips blank records).  Anything else is lower-cased with Python's Unicode rules
    per record, blank records (str.isspace semantics of memo_cli.py:136-142) are dropped here, and the ids of the
    kept records are returned explicitly."""
    texts = ["" if t is None else t for t in texts]
    n = len(texts)
    joined = "\n".join(texts)
    if joined.isascii():
        blob = joined.encode("ascii")
        offsets = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter(map(len, texts), dtype=np.int64, count=n) + 1, out=offsets[1:])
        offsets[n] = len(blob)  # the last record has no trailing "\n"; the separators inside the ranges are white space
        return blob, offsets, None
    base_ids = np.arange(n, dtype=np.int64) if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
    keep = [i for i, t in enumerate(texts) if not (skip_blank and (t == "" or t.isspace()))]
    low = [texts[i].lower().encode("utf-8") for i in keep]
    offsets = np.zeros(len(keep) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in low], out=offsets[1:])
    return b"".join(low), offsets, base_ids[keep]


class Index:
    """Common base (faiss.Index)."""

    d: int
    metric_type: int
    is_trained = True
    verbose = False

    @property
    def ntotal(self) -> int:
        raise NotImplementedError

    def train(self, x) -> None:  # flat indexes need no training [upstream]
        return None


class IndexFlat(Index):
    """Exhaustive index resident in HBM.  `store` = "f32" | "bf16"; `normalize` L2-normalises rows
    at add time and queries at search time (cosine as IP over unit vectors, memo_cli.py:131-135)."""

    def __init__(self, d: int, metric: int = METRIC_L2, *, store: str = "f32", normalize: bool = False,
                 device: int | None = None):
        self._h = C.c_void_p()
        self.d = int(d)
        self.metric_type = int(metric)
        self.store = store
        self.normalize = bool(normalize)
        self.device = _default_device() if device is None else int(device)
        st = {"f32": _cabi.STORE_F32, "fp32": _cabi.STORE_F32, "bf16": _cabi.STORE_BF16}.get(store)
        if st is None:
            raise ValueError(f"unknown store {store!r}")
        L = _cabi.load()
        _cabi.check(L.b200_index_create(C.byref(self._h), self.d, self.metric_type, st, self.device))
        if self.normalize:
            self.set_option("normalize_queries", 1)

    # -- lifetime --
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), C.c_void_p()
        if h is not None and h.value:
            _cabi.load().b200_index_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options (sweep harness / bench) --
    def set_option(self, name: str, value: int) -> None:
        _cabi.check(_cabi.load().b200_index_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        out = C.c_int64(0)
        _cabi.check(_cabi.load().b200_index_get_option(self._h, name.encode(), C.byref(out)))
        return out.value

    # -- faiss surface --
    @property
    def ntotal(self) -> int:
        return int(_cabi.load().b200_index_ntotal(self._h))

    @property
    def launch_count(self) -> int:
        return int(_cabi.load().b200_index_launch_count(self._h))

    def reserve(self, n_total: int) -> None:
        _cabi.check(_cabi.load().b200_index_reserve(self._h, int(n_total)))

    def reset(self) -> None:
        _cabi.check(_cabi.load().b200_index_reset(self._h))

    def _coerce_x(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2, "expected a 2-D array [n, d]"
        assert x.shape[1] == self.d, f"vector dimension {x.shape[1]} != index dimension {self.d}"
        return x

    def add(self, x) -> None:
        x = self._coerce_x(x)
        _cabi.check(_cabi.load().b200_index_add(self._h, x.ctypes.data, x.shape[0], None, int(self.normalize)))

    def add_with_ids(self, x, ids) -> None:
        raise RuntimeError("add_with_ids not implemented for this type of index")  # as faiss [upstream]

    def add_device(self, x, ids=None) -> None:
        """Rows (and optional int64 ids) already resident on this index's GPU as torch tensors:
        K1 ingests them without touching the host (b200_index_add_dev)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[1] == self.d
        idp = None
        if ids is not None:
            assert ids.is_cuda and ids.dtype == torch.int64 and ids.is_contiguous() and ids.shape == (x.shape[0],)
            idp = ids.data_ptr()
        torch.cuda.current_stream(x.device).synchronize()  # the handle's stream must see finished producers
        _cabi.check(_cabi.load().b200_index_add_dev(self._h, x.data_ptr(), x.shape[0], idp, int(self.normalize)))

    def _add_with_ids(self, x, ids) -> None:
        x = self._coerce_x(x)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        assert ids.shape == (x.shape[0],), "not same number of vectors and ids"
        _cabi.check(_cabi.load().b200_index_add(self._h, x.ctypes.data, x.shape[0], ids.ctypes.data, int(self.normalize)))

    def _add_texts(self, blob: bytes, offsets: np.ndarray, ids: np.ndarray | None, first_id: int, skip_blank: bool,
                   with_ids: bool) -> int:
        """K6 (b200_index_add_texts): records = byte ranges of `blob`; embedded, normalised and stored on the device.
        Returns the number of rows added."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.shape[0] - 1
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            assert ids.shape == (n,), "not same number of texts and ids"
            idp = ids.ctypes.data
        added = C.c_int64(0)
        _cabi.check(_cabi.load().b200_index_add_texts(self._h, blob, offsets.ctypes.data, n, idp, int(first_id), int(skip_blank),
                                                      1, int(with_ids), C.byref(added)))
        return int(added.value)

    def add_synthetic(self, n: int, seed: int, first_row: int = 0, *, with_ids: bool = False, first_id: int = 0) -> None:
        """Rows u(seed,row,col) generated on the device; for databases too large to upload."""
        _cabi.check(_cabi.load().b200_index_add_synthetic(self._h, int(n), int(seed), int(first_row),
                                                          int(self.normalize), int(with_ids), int(first_id)))

    def search(self, x, k: int, row_mask=None, ids_allowed=None):
        """faiss Index.search.  Extensions, the filter runs inside the search kernels:
        `row_mask` — bool array over row positions, only rows whose entry is True can be returned;
        `ids_allowed` — iterable of ids that may be returned (the row bitmap is built on the device)."""
        x = self._coerce_x(x)
        k = int(k)
        assert k > 0
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        if ids_allowed is not None:
            if row_mask is not None:
                raise ValueError("give row_mask or ids_allowed, not both")
            if isinstance(ids_allowed, np.ndarray):
                allowed = np.ascontiguousarray(ids_allowed, dtype=np.int64).reshape(-1)
            else:
                allowed = np.fromiter((int(i) for i in ids_allowed), dtype=np.int64)
            _cabi.check(_cabi.load().b200_index_search_ids_allowed(
                self._h, x.ctypes.data, nq, k, allowed.ctypes.data if allowed.size else None, allowed.size,
                D.ctypes.data, I.ctypes.data))
        elif row_mask is None:
            _cabi.check(_cabi.load().b200_index_search(self._h, x.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data))
        else:
            bits = pack_row_mask(row_mask, self.ntotal)
            _cabi.check(_cabi.load().b200_index_search_masked(self._h, x.ctypes.data, nq, k, bits.ctypes.data,
                                                              D.ctypes.data, I.ctypes.data))
        return D, I

    def search_device(self, q, k: int, D=None, I=None, stream: int | None = None):
        """Device-resident search: q/D/I are torch CUDA tensors on this index's device; the work is
        enqueued on `stream` (a raw cudaStream_t, default torch's current stream), not synchronised."""
        import torch

        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous() and q.dim() == 2 and q.shape[1] == self.d
        nq = q.shape[0]
        if D is None:
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        if I is None:
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        if stream is None:
            stream = torch.cuda.current_stream(q.device).cuda_stream
        if not stream:
            stream = 1  # cudaStreamLegacy: torch's default stream (NULL would mean the handle's own stream)
        _cabi.check(_cabi.load().b200_index_search_dev(self._h, q.data_ptr(), nq, int(k), D.data_ptr(), I.data_ptr(),
                                                       C.c_void_p(stream)))
        return D, I

    def sync(self) -> None:
        _cabi.check(_cabi.load().b200_index_sync(self._h))

    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else int(n)
        out = np.empty((n, self.d), dtype=np.float32)
        if n:
            _cabi.check(_cabi.load().b200_index_get_rows(self._h, int(i0), n, out.ctypes.data))
        return out

    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(int(i), 1)[0]

    def _ids(self) -> np.ndarray:
        out = np.empty((self.ntotal,), dtype=np.int64)
        if out.size:
            _cabi.check(_cabi.load().b200_index_get_ids(self._h, out.ctypes.data))
        return out


class IndexFlatIP(IndexFlat):
    def __init__(self, d: int, **kw):
        super().__init__(d, METRIC_INNER_PRODUCT, **kw)


class IndexFlatL2(IndexFlat):
    def __init__(self, d: int, **kw):
        super().__init__(d, METRIC_L2, **kw)


class IndexHNSWFlat(IndexFlat):
    """memo constructs IndexHNSWFlat(384, 32) and sets hnsw.efConstruction / hnsw.efSearch
    (memo_cli.py:245-247) but always searches with k = ntotal (:291), i.e. exhaustively.  This class
    accepts the same constructor and attribute writes and is an exact flat index."""

    def __init__(self, d: int, M: int = 32, metric: int = METRIC_L2, **kw):
        super().__init__(d, metric, **kw)
        self.hnsw = SimpleNamespace(efConstruction=40, efSearch=16, M=int(M))


class IndexIDMap(Index):
    """faiss.IndexIDMap: user ids over a base index; ids are translated inside the search kernel."""

    _fourcc = b"IxMp"

    def __init__(self, index: IndexFlat):
        if not isinstance(index, IndexFlat):
            raise RuntimeError("IndexIDMap: only flat base indexes are supported")
        if index.ntotal != 0:
            raise RuntimeError("index must be empty on input")  # as faiss [upstream]
        self.index = index
        self.own_fields = True
        self.d = index.d
        self.metric_type = index.metric_type

    @property
    def ntotal(self) -> int:
        return self.index.ntotal

    @property
    def id_map(self) -> Int64Vector:
        return Int64Vector(self.index._ids())

    def add_with_ids(self, x, ids) -> None:
        self.index._add_with_ids(x, ids)

    def add_texts(self, texts, ids=None, *, skip_blank: bool = True) -> int:
        """Extension: embed_text_hash + normalize + add_with_ids of every record on the device — the
        whole of rebuild_index_from_texts (memo_cli.py:272-285) in one call.  ids default to the records' positions in
        `texts` (what the reference assigns, :276-282); blank records are skipped as there.  Token hashes are CPython's
        under PYTHONHASHSEED=0 (reproducible across processes).  Returns the number of rows added."""
        blob, offsets, keep_ids = pack_texts(texts, ids, skip_blank)
        if keep_ids is None:  # plain ASCII: blank detection and ASCII lower-casing happen on the device
            return self.index._add_texts(blob, offsets, ids, 0, skip_blank, True)
        return self.index._add_texts(blob, offsets, keep_ids, 0, False, True)

    def add(self, x) -> None:
        raise RuntimeError("add does not make sense with IndexIDMap, use add_with_ids")  # as faiss

    def search(self, x, k: int, row_mask=None, ids_allowed=None):
        """`ids_allowed` (extension): iterable of record ids that may be returned (filter push-down)."""
        return self.index.search(x, k, row_mask=row_mask, ids_allowed=ids_allowed)

    def search_device(self, q, k: int, **kw):
        return self.index.search_device(q, k, **kw)

    def reset(self) -> None:
        self.index.reset()

    def reserve(self, n_total: int) -> None:
        self.index.reserve(n_total)


class IndexIDMap2(IndexIDMap):
    """IndexIDMap plus reconstruct-by-id (reverse map built on demand)."""

    _fourcc = b"IxM2"

    def reconstruct(self, key: int) -> np.ndarray:
        ids = self.index._ids()
        pos = np.nonzero(ids == int(key))[0]
        if pos.size == 0:
            raise RuntimeError(f"key {key} not found")
        return self.index.reconstruct(int(pos[-1]))


# ------------------------------------------------------------------------------------------------
# normalize_L2 — K1 as a function
# ------------------------------------------------------------------------------------------------
def normalize_L2(x: np.ndarray, device: int | None = None) -> None:
    """In-place row normalisation on the device with memo's semantics (memo_cli.py:131-135)."""
    assert isinstance(x, np.ndarray) and x.dtype == np.float32 and x.ndim == 2 and x.flags.c_contiguous
    dev = _default_device() if device is None else device
    _cabi.check(_cabi.load().b200_normalize_rows(x.ctypes.data, x.shape[0], x.shape[1], dev))


# ------------------------------------------------------------------------------------------------
# .memo files — faiss's native index serialisation [upstream layout]
# ------------------------------------------------------------------------------------------------
def write_index(index: Index, path: str) -> None:
    """faiss.write_index (memo_cli.py:361, :448).  bf16-stored rows are widened to fp32 (lossless).
    The library writes the faiss headers (b200_memo_write_headers) and moves the rows and ids
    device -> pinned ring -> file (b200_index_write_file); nothing passes through numpy."""
    path = os.fspath(path)
    if isinstance(index, IndexIDMap):
        base, kind = index.index, (2 if isinstance(index, IndexIDMap2) else 1)
    elif isinstance(index, IndexFlat):
        base, kind = index, 0
    else:
        raise RuntimeError(f"don't know how to serialize {type(index).__name__}")
    L = _cabi.load()
    info = _cabi.MemoInfo(kind=kind, d=base.d, metric=base.metric_type, ntotal=base.ntotal)
    rows_off, ids_off = C.c_int64(0), C.c_int64(-1)
    _cabi.check(L.b200_memo_write_headers(path.encode(), C.byref(info), C.byref(rows_off), C.byref(ids_off)))
    if info.ntotal:
        _cabi.check(L.b200_index_write_file(base._h, path.encode(), rows_off.value, ids_off.value))


def read_index(path: str, device: int | None = None) -> Index:
    """faiss.read_index (memo_cli.py:255): raises on a missing or corrupt file (memo catches
    Exception and starts a fresh index, :256-257).  The library parses the headers (b200_memo_probe:
    IxM2 / IxMp wrappers, flat payloads, memo's own IHNf files with the graph skipped) and streams the
    payload file -> pinned ring -> device (b200_index_add_file)."""
    path = os.fspath(path)
    os.stat(path)  # FileNotFoundError for a missing file, like open()
    L = _cabi.load()
    info = _cabi.MemoInfo()
    _cabi.check(L.b200_memo_probe(path.encode(), C.byref(info)))
    base = IndexFlat(info.d, info.metric, device=device)
    out = base if info.kind == 0 else (IndexIDMap2 if info.kind == 2 else IndexIDMap)(base)
    if info.ntotal:
        _cabi.check(L.b200_index_add_file(base._h, path.encode(), info.rows_offset, info.ntotal, info.ids_offset, 0))
    return out
