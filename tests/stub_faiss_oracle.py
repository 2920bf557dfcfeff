"""A `faiss` stand-in backed by the CPU ORACLE.  Test infrastructure only: it lets the reference's
own memo_cli.py be imported in the build container (faiss-cpu is not installable there) so that
tests/golden/make_golden.py can record what the REFERENCE's functions produce on top of exact flat
semantics.  The product package never imports this."""
from __future__ import annotations

import pickle
from types import SimpleNamespace

import numpy as np

from oracle import oracle

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


class _Vec:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.int64)


def vector_to_array(v):
    return v.a.copy()


class IndexFlat:
    def __init__(self, d, metric=METRIC_L2):
        self.d, self.metric_type = d, metric
        self.rows = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.rows.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.shape[1] == self.d
        self.rows = np.concatenate([self.rows, x], axis=0)

    def search(self, x, k, ids=None):
        # STUB_FAISS_ORDER=device: the B200 kernels' summation order (bit-exact transcripts against the GPU shim);
        # default: the faiss-like SIMD order
        import os

        order = oracle.ORDER_DEVICE if os.environ.get("STUB_FAISS_ORDER") == "device" else oracle.ORDER_SIMD
        return oracle.search(self.metric_type, self.rows, x, k, ids=ids, order=order)


class IndexFlatIP(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_INNER_PRODUCT)


class IndexFlatL2(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_L2)


class IndexHNSWFlat(IndexFlat):
    def __init__(self, d, M=32):
        super().__init__(d, METRIC_L2)  # faiss default metric for HNSWFlat [upstream]
        self.hnsw = SimpleNamespace(efConstruction=40, efSearch=16)


class IndexIDMap2:
    def __init__(self, base):
        self.index = base
        self.d = base.d
        self.ids = np.zeros((0,), dtype=np.int64)

    @property
    def ntotal(self):
        return self.index.ntotal

    @property
    def id_map(self):
        return _Vec(self.ids)

    def add_with_ids(self, x, ids):
        self.index.add(x)
        self.ids = np.concatenate([self.ids, np.asarray(ids, dtype=np.int64)])

    def search(self, x, k):
        return self.index.search(x, k, ids=self.ids)


IndexIDMap = IndexIDMap2


def write_index(index, path):
    with open(path, "wb") as f:
        pickle.dump(index, f)


def read_index(path):
    with open(path, "rb") as f:
        return pickle.load(f)
