#!/usr/bin/env python3
"""Rebuild from texts (BASELINE config 4: "full rebuild (add) from YAML-derived embeddings"; memo_cli.py:272-285).

Times, for a synthetic corpus of N records (~60 tokens / ~400 bytes each, built from a 100k-record pool):
  c_abi      b200_index_add_texts(host blob): pinned staging + H2D of the text + tokenise/hash/bucket/normalise/store
             kernels — text bytes in host memory -> resident rows (the number a non-Python host sees)
  python     memo_adapter.rebuild_index_from_texts(list[str], hash_fn=None): the same plus Python's join/encode
  host_embed the round-1 path on a sample: b200_hash_embed on the host cores + host-fed add (H2D of fp32 rows + K1)
One JSON line per (what, d, store).
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import c99_vectordb_b200 as m
from c99_vectordb_b200 import _cabi
from c99_vectordb_b200 import memo_adapter as ma


def pool(n, seed=0):
    rng = np.random.default_rng(seed)
    vocab = [f"w{i}" for i in range(5000)] + ["peanuts", "allergy", "wifi", "password", "rotate", "API", "keys", "Sarah"]
    out = []
    for _ in range(n):
        k = int(rng.integers(40, 80))
        out.append(" ".join(vocab[int(j)] for j in rng.integers(0, len(vocab), size=k)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--py-n", type=int, default=1_000_000)
    ap.add_argument("--host-n", type=int, default=500_000)
    a = ap.parse_args()
    recs = pool(100_000)
    enc = [(r + "\n").encode("ascii") for r in recs]
    lens = np.fromiter((len(b) for b in enc), dtype=np.int64, count=len(enc))
    reps = -(-a.n // len(enc))
    blob = b"".join(enc) * reps
    offsets = np.zeros(reps * len(enc) + 1, dtype=np.int64)
    np.cumsum(np.tile(lens, reps), out=offsets[1:])
    offsets = offsets[: a.n + 1].copy()
    nbytes = int(offsets[-1])
    L = _cabi.load()
    for d, store in ((384, "f32"), (1024, "bf16")):
        best = None
        for rep in range(3):
            idx = m.IndexIDMap2(m.IndexFlat(d, 1, store=store))
            idx.reserve(a.n)
            added = C.c_int64(0)
            t0 = time.perf_counter()
            _cabi.check(L.b200_index_add_texts(idx.index._h, blob, offsets.ctypes.data, a.n, None, 0, 1, 1, 1, C.byref(added)))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert added.value == a.n
            launches = idx.index.launch_count
            idx.index.close()
        print(json.dumps({"what": "c_abi add_texts (host text bytes -> resident rows)", "records": a.n, "d": d, "store": store,
                          "text_GB": round(nbytes / 1e9, 3), "seconds": round(best, 4), "records_per_s": round(a.n / best),
                          "text_GBps": round(nbytes / best / 1e9, 2), "launches": launches}), flush=True)
        texts = (recs * (-(-a.py_n // len(recs))))[: a.py_n]
        best = None
        for rep in range(2):
            t0 = time.perf_counter()
            idx = ma.rebuild_index_from_texts(texts, dim=d, hash_fn=None) if store == "f32" else None
            if idx is None:
                idx = m.IndexIDMap2(m.IndexFlat(d, 1, store=store))
                idx.add_texts(texts)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert idx.ntotal == a.py_n
            idx.index.close()
        print(json.dumps({"what": "python rebuild_index_from_texts(list[str]) incl. join/encode", "records": a.py_n, "d": d, "store": store,
                          "seconds": round(best, 4), "records_per_s": round(a.py_n / best)}), flush=True)
        texts = texts[: a.host_n]
        t0 = time.perf_counter()
        rows = ma.embed_texts_stable(texts, d)
        t1 = time.perf_counter()
        idx = m.IndexIDMap2(m.IndexFlat(d, 1, store=store, normalize=True))
        idx.add_with_ids(rows, np.arange(len(texts), dtype=np.int64))
        t2 = time.perf_counter()
        idx.index.close()
        print(json.dumps({"what": "round-1 path: host embedder + host-fed add", "records": len(texts), "d": d, "store": store,
                          "embed_s": round(t1 - t0, 4), "add_s": round(t2 - t1, 4), "records_per_s": round(len(texts) / (t2 - t0))}), flush=True)


if __name__ == "__main__":
    main()
