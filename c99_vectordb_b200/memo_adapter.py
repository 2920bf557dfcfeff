"""Host-side mirror of memo's embedding + index adapter layer (memo_cli.py:131-167, :244-298).

Same names, argument meaning and results as the reference functions, restated for a resident GPU
index: a rebuild embeds every record on the host and issues ONE add_with_ids([N,d], ids[N])
instead of N single-row crossings (memo_cli.py:276-282), and search_all can ask for a bounded k.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Iterable, Sequence

import numpy as np

from . import index as _ix

DIM = 384    # memo_cli.py:17
MAX_K = 100  # memo_cli.py:18

_TOKEN = re.compile(r"[a-zA-Z0-9_]+")
_SPACE = re.compile(r"\s+")


@dataclass
class Result:  # memo_cli.py:21-24
    doc_id: int
    score: float


def normalize_whitespace(text: str) -> str:  # memo_cli.py:138-139
    return _SPACE.sub(" ", text).strip()


def is_blank_body(text: str | None) -> bool:  # memo_cli.py:142-143
    return text is None or normalize_whitespace(text) == ""


def bag_of_hashed_words(text: str, dim: int = DIM, hash_fn: Callable[[str], int] = hash) -> np.ndarray:
    """The un-normalised signed hashing-trick vector of memo_cli.py:158-166.  `hash_fn` defaults to
    Python's builtin hash() exactly as the reference (salted per process unless PYTHONHASHSEED is
    fixed — SURVEY.md §0.4)."""
    vec = np.zeros((dim,), dtype=np.float32)
    for tok in _TOKEN.findall(normalize_whitespace(text).lower()):
        h = hash_fn(tok)
        vec[abs(h) % dim] += 1.0 if (h & 1) else -1.0
    return vec


def stable_hash(token: str) -> int:
    """CPython's hash(str) with the PYTHONHASHSEED=0 key (SipHash-1-3, zero key): the same in every
    process, unlike the builtin the reference calls (memo_cli.py:163)."""
    from . import _cabi

    b = token.encode("utf-8")
    return int(_cabi.load().b200_py_hash_seed0(b, len(b)))


def embed_texts_stable(texts: Sequence[str], dim: int = DIM) -> np.ndarray:
    """Native bulk embedder: [n, dim] un-normalised rows, bit-identical to the reference's
    embed_text_hash buckets under PYTHONHASHSEED=0.  One C call instead of a Python token loop."""
    from . import _cabi

    n = len(texts)
    offsets = np.zeros(n + 1, dtype=np.int64)
    joined = "\n".join(texts)
    if joined.isascii():
        # one lower() / encode() over everything: ASCII lower-casing keeps lengths, and the "\n" that ends up inside
        # each record's byte range is not a token byte
        blob = joined.lower().encode("ascii")
        np.cumsum(np.fromiter(map(len, texts), dtype=np.int64, count=n) + 1, out=offsets[1:])
        offsets[n] = len(blob)
    else:
        low = [t.lower().encode("utf-8") for t in texts]  # str.lower() stays in Python (Unicode case rules)
        np.cumsum([len(b) for b in low], out=offsets[1:])
        blob = b"".join(low)
    out = np.empty((n, dim), dtype=np.float32)
    rc = _cabi.load().b200_hash_embed(blob, offsets.ctypes.data, n, dim, out.ctypes.data)
    if rc:
        raise RuntimeError("b200_hash_embed failed")
    return out


def embed_texts(texts: Sequence[str], dim: int = DIM, hash_fn: Callable[[str], int] | None = hash) -> np.ndarray:
    """[n, dim] un-normalised rows; normalisation happens on the device at add / search time (K1).
    hash_fn=None selects the native stable-hash embedder."""
    if hash_fn is None:
        return embed_texts_stable(texts, dim)
    out = np.zeros((len(texts), dim), dtype=np.float32)
    for i, t in enumerate(texts):
        out[i] = bag_of_hashed_words(t, dim, hash_fn)
    return out


def normalize(v: np.ndarray) -> np.ndarray:
    """memo_cli.py:131-135 on the device (K1): zero vector when the norm is <= 1e-8."""
    x = np.array(v, dtype=np.float32, copy=True).reshape(1, -1)
    _ix.normalize_L2(x)
    return x.reshape(np.shape(v))


def embed_text_hash(text: str, dim: int = DIM, hash_fn: Callable[[str], int] = hash) -> np.ndarray:
    """memo_cli.py:158-167: unit-norm fp32[dim]."""
    return normalize(bag_of_hashed_words(text, dim, hash_fn))


def create_index(dim: int = DIM, metric: str = "l2", device: int | None = None) -> _ix.IndexIDMap2:
    """memo_cli.py:244-248 restated as IndexIDMap2 over an exact flat index (north_star).  "l2" keeps
    the reference's default metric (scores are squared L2, ascending); "ip" is cosine on unit rows."""
    base = _ix.IndexFlatL2(dim, device=device) if metric == "l2" else _ix.IndexFlatIP(dim, device=device)
    return _ix.IndexIDMap2(base)


def load_index(path: Path, verbose: bool = False, dim: int = DIM, metric: str = "l2") -> _ix.IndexIDMap2:
    """memo_cli.py:251-262: missing or unreadable file -> fresh index; bare index -> wrapped."""
    if not Path(path).exists():
        return create_index(dim, metric)
    try:
        idx = _ix.read_index(str(path))
    except Exception:
        return create_index(dim, metric)
    if isinstance(idx, _ix.IndexIDMap2):
        return idx
    if isinstance(idx, _ix.IndexIDMap):
        idx.__class__ = _ix.IndexIDMap2
        return idx
    rows = idx.reconstruct_n(0, idx.ntotal)
    wrapped = create_index(idx.d, "ip" if idx.metric_type == _ix.METRIC_INNER_PRODUCT else "l2")
    if rows.shape[0]:
        wrapped.add_with_ids(rows, np.arange(rows.shape[0], dtype=np.int64))
    return wrapped


def get_existing_ids(index: _ix.IndexIDMap2) -> set[int]:  # memo_cli.py:265-269
    if index.ntotal == 0:
        return set()
    return set(int(x) for x in _ix.vector_to_array(index.id_map).tolist())


def rebuild_index_from_texts(texts: Iterable[str | None], verbose: bool = False, dim: int = DIM,
                             metric: str = "l2", hash_fn: Callable[[str], int] = hash,
                             vectors: np.ndarray | None = None) -> _ix.IndexIDMap2:
    """memo_cli.py:272-285: full rebuild; blank records are skipped so ids may be sparse.  One bulk
    add.  `vectors` (unit rows for the kept records, in order) bypasses the host embedder."""
    texts = list(texts)
    idx = create_index(dim, metric)
    if hash_fn is None and vectors is None:
        # stable hash: the whole rebuild — tokenise, hash, bucket, normalise, add — is one device pipeline (K6);
        # only the text bytes cross PCIe
        idx.add_texts(texts)
        return idx
    keep = [i for i, t in enumerate(texts) if not is_blank_body(t or "")]
    if keep:
        if vectors is None:
            rows = embed_texts([texts[i] or "" for i in keep], dim, hash_fn)
            _ix.normalize_L2(rows)
        else:
            rows = np.ascontiguousarray(vectors, dtype=np.float32)
            assert rows.shape == (len(keep), dim)
        idx.add_with_ids(rows, np.asarray(keep, dtype=np.int64))
    return idx


def search_all(index: _ix.IndexIDMap2, query_vec: np.ndarray, k: int | None = None) -> list[Result]:
    """memo_cli.py:288-298: the full best-first ranking (k = ntotal) with id < 0 entries dropped.
    Passing k bounds the work to the fused top-k kernel."""
    if index.ntotal == 0:
        return []
    kk = int(index.ntotal) if k is None else int(k)
    scores, ids = index.search(np.asarray(query_vec, dtype=np.float32).reshape(1, -1), kk)
    return [Result(int(i), float(s)) for s, i in zip(scores[0].tolist(), ids[0].tolist()) if i >= 0]


def search_filtered(index: _ix.IndexIDMap2, query_vec: np.ndarray, k: int, allowed_ids: Iterable[int]) -> list[Result]:
    """Filter first, then rank (what SKILL.md:243-247 documents and memo_cli.py:479-506 does the other
    way round): the metadata predicate is evaluated on the host, becomes a row bitmap, and the scan
    kernel returns the best k among the allowed records only — no k = ntotal ranking, no O(N) Python
    loop."""
    if index.ntotal == 0:
        return []
    scores, ids = index.search(np.asarray(query_vec, dtype=np.float32).reshape(1, -1), int(k), ids_allowed=allowed_ids)
    return [Result(int(i), float(s)) for s, i in zip(scores[0].tolist(), ids[0].tolist()) if i >= 0]


def recall(index: _ix.IndexIDMap2, texts: Sequence[str | None], metas: Sequence[dict | None], query_vec: np.ndarray,
           k: int, predicate: Callable[[dict], bool] | None = None) -> list[Result]:
    """The selection loop of command_recall (memo_cli.py:491-521) with the filter pushed down: a record
    is eligible iff its id indexes `texts`, its body is not blank and (when a predicate is given) it has
    non-empty metadata that satisfies the predicate; the first k eligible results best-first, dropping
    scores below -0.9 (:494)."""
    eligible = []
    for doc_id, text in enumerate(texts):
        if is_blank_body(text or ""):
            continue
        if predicate is not None:
            rec = metas[doc_id] if doc_id < len(metas) and metas[doc_id] is not None else {}
            if not rec or not predicate(rec):
                continue
        eligible.append(doc_id)
    out = search_filtered(index, query_vec, k, eligible)
    return [r for r in out if not (r.score < -0.9)][:k]
