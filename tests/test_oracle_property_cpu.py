"""Property tests of the oracle on integer-valued vectors: every product and every partial sum is exact in fp32
whatever the summation order, so scores are plain integers, ties are exact and plentiful, and the expected answer
can be written down with integer arithmetic alone — (score best-first, smaller row position first), id -1 and
-/+FLT_MAX padding past the database.  Every summation order of the oracle must agree
with it bit for bit."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle

FLT_MAX = np.finfo(np.float32).max


def brute_force(metric, db, q, k, ids=None):
    """Integer arithmetic + a stable sort on (score, row): the tie rule by construction."""
    n = db.shape[0]
    D = np.full((q.shape[0], k), -FLT_MAX if metric == 0 else FLT_MAX, np.float32)
    I = np.full((q.shape[0], k), -1, np.int64)
    dbi, qi = db.astype(np.int64), q.astype(np.int64)
    for j in range(q.shape[0]):
        s = dbi @ qi[j] if metric == 0 else ((dbi - qi[j]) ** 2).sum(axis=1)
        order = np.argsort(-s if metric == 0 else s, kind="stable")[:k]
        m = order.shape[0]
        D[j, :m] = s[order].astype(np.float32)
        I[j, :m] = order if ids is None else ids[order]
    return D, I


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 300), d=st.sampled_from([1, 3, 8, 33, 100, 384]), k=st.integers(1, 40), nq=st.integers(1, 4),
       metric=st.sampled_from([0, 1]), span=st.integers(0, 3), with_ids=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_oracle_equals_integer_brute_force(n, d, k, nq, metric, span, with_ids, seed):
    rng = np.random.default_rng(seed)
    db = rng.integers(-span, span + 1, size=(n, d)).astype(np.float32)
    q = rng.integers(-span, span + 1, size=(nq, d)).astype(np.float32)
    ids = (rng.permutation(10 * n)[:n].astype(np.int64) - 3 * n) if with_ids else None  # sparse, some negative
    Dw, Iw = brute_force(metric, db, q, k, ids)
    for order in (oracle.ORDER_SIMD, oracle.ORDER_DEVICE):
        D, I = oracle.search(metric, db, q, k, ids=ids, order=order)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    Dr, Ir = oracle.search(metric, db, q, k, ids=ids, rowpar=True)
    np.testing.assert_array_equal(Ir, Iw)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(2, 200), G=st.integers(2, 5), k=st.integers(1, 30), metric=st.sampled_from([0, 1]), seed=st.integers(0, 2**31 - 1))
def test_sharded_merge_of_tied_lists_equals_unsharded(n, G, k, metric, seed):
    """Row shards + merge (rank-major, then list position) reproduce the global tie rule even when most scores tie."""
    rng = np.random.default_rng(seed)
    d = 16
    db = rng.integers(-1, 2, size=(n, d)).astype(np.float32)
    q = rng.integers(-1, 2, size=(2, d)).astype(np.float32)
    ids = np.arange(n, dtype=np.int64) * 5 + 1
    per = -(-n // G)
    parts = []
    for g in range(G):
        lo, hi = min(n, g * per), min(n, (g + 1) * per)
        parts.append(oracle.search(metric, db[lo:hi], q, k, ids=ids[lo:hi]) if hi > lo else
                     (np.full((2, k), -FLT_MAX if metric == 0 else FLT_MAX, np.float32), np.full((2, k), -1, np.int64)))
    D, I = oracle.merge_topk(metric, np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    Dw, Iw = brute_force(metric, db, q, k, ids)
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)
