"""Oracle-independent parity: on integer-valued vectors every product and partial sum is exact in fp32 (and the small
integers survive bf16), so the expected ids and distances follow from integer arithmetic and a stable sort alone — no
restatement of the kernel's summation order is involved.  Scores tie by the thousand, which makes this the hardest test
of the stated tie rule (score best-first, then smaller row position): every path must reproduce it exactly — the scan
kernel, the full ranking, both tensor-core forms with their certificates and fallbacks, the streamed shadow, the
pre-filter, filtered search and row shards."""
import ctypes as C

import numpy as np
import pytest

from test_oracle_property_cpu import brute_force

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


def make(n, d, nq, span, seed):
    rng = np.random.default_rng(seed)
    db = rng.integers(-span, span + 1, size=(n, d)).astype(np.float32)
    q = rng.integers(-span, span + 1, size=(nq, d)).astype(np.float32)
    return db, q


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("n,d,span", [(30_000, 32, 1), (12_000, 100, 2), (9_001, 384, 1)])
def test_scan_and_full_ranking(b200, metric, n, d, span):
    db, q = make(n, d, 3, span, n + d)
    ids = np.arange(n, dtype=np.int64) * 3 - 7
    idx = b200.IndexIDMap2(b200.IndexFlat(d, metric))
    idx.add_with_ids(db, ids)
    idx.index.set_option("gemm_min_nq", 0)  # scan kernel (query blocks) only
    for k in (1, 10, 100, 256):
        D, I = idx.search(q, k)
        Dw, Iw = brute_force(metric, db, q, k, ids)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    D, I = idx.search(q[:1], n)  # memo's call: k = ntotal, full ranking
    Dw, Iw = brute_force(metric, db, q[:1], n, ids)
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("nq,rows_form", [(40, 1), (40, 0), (300, 1)])
def test_tensor_core_forms_under_massive_ties(b200, metric, nq, rows_form):
    """Thousands of rows share the k-th score: the certificate cannot separate them and must hand over to the widened
    pass / the exact scan; whatever path answers, the tie rule decides."""
    n, d, k = 70_000, 128, 10
    db, q = make(n, d, nq, 1, 77 + nq)
    idx = b200.IndexFlat(d, metric)
    idx.add(db)
    idx.set_option("gemm_rows_form", rows_form)
    D, I = idx.search(q, k)
    assert idx.get_option("stat_gemm_used") == 1
    Dw, Iw = brute_force(metric, db, q, k)
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)
    idx.set_option("gemm_shadow_max_rows", 32_768)  # streamed shadow
    D, I = idx.search(q, k)
    assert idx.get_option("stat_gemm_streamed") == 1
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)


@pytest.mark.parametrize("metric", [0, 1])
def test_prefilter_and_filtered_search(b200, metric):
    n, d, k = 80_000, 256, 10
    db, q = make(n, d, 4, 2, 5)
    idx = b200.IndexFlat(d, metric)
    idx.add(db)
    idx.set_option("prefilter", 1)
    for i in range(q.shape[0]):
        D, I = idx.search(q[i:i + 1], k)
        Dw, Iw = brute_force(metric, db, q[i:i + 1], k)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    assert idx.get_option("stat_prefilter_used") == 1
    idx.set_option("prefilter", 0)
    mask = np.random.default_rng(1).random(n) < 0.2
    rows = np.nonzero(mask)[0]
    for qq in (q[:1], q):  # scan path and tensor-core path
        D, I = idx.search(qq, k, row_mask=mask)
        Dw, Iw = brute_force(metric, db[rows], qq, k, rows.astype(np.int64))
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)


@pytest.mark.parametrize("metric", [0, 1])
def test_row_shards_merge_with_the_global_tie_rule(b200, metric):
    """G shards searched one after the other on this GPU (b200_index_search_shard_dev), merged and certified by
    b200_merge_certify_dev; uncertified queries repeat with widen = 1, 2 — as sharded.py does across ranks."""
    import torch

    from c99_vectordb_b200 import _cabi
    from c99_vectordb_b200.sharded import shard_range

    L = _cabi.load()
    n, d, k, nq, world = 60_000, 64, 20, 24, 3
    db, q = make(n, d, nq, 1, 11)
    ids = np.arange(n, dtype=np.int64)
    shards = []
    for g in range(world):
        lo, hi = shard_range(n, world, g)
        ix = b200.IndexIDMap2(b200.IndexFlat(d, metric))
        ix.add_with_ids(db[lo:hi], ids[lo:hi])
        shards.append(ix)
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
    Dw, Iw = brute_force(metric, db, q, k, ids)
    todo = np.arange(nq)
    D_all = np.zeros((nq, k), np.float32)
    I_all = np.zeros((nq, k), np.int64)
    for widen in (0, 1, 2):
        m = todo.size
        qd = torch.from_numpy(q[todo]).to(dev)
        Dp = torch.empty((world, m, k), dtype=torch.float32, device=dev)
        Ip = torch.empty((world, m, k), dtype=torch.int64, device=dev)
        Bp = torch.empty((world, m), dtype=torch.float32, device=dev)
        for g, ix in enumerate(shards):
            _cabi.check(L.b200_index_search_shard_dev(ix.index._h, qd.data_ptr(), m, k, world, widen, Dp[g].data_ptr(), Ip[g].data_ptr(),
                                                      Bp[g].data_ptr(), st))
        D = torch.empty((m, k), dtype=torch.float32, device=dev)
        I = torch.empty((m, k), dtype=torch.int64, device=dev)
        unc = torch.zeros(m, dtype=torch.int32, device=dev)
        n_unc = torch.zeros(1, dtype=torch.int32, device=dev)
        _cabi.check(L.b200_merge_certify_dev(metric, world, m, k, n, Dp.data_ptr(), Ip.data_ptr(), 0, 0, Bp.data_ptr(), 0, D.data_ptr(),
                                             I.data_ptr(), unc.data_ptr(), n_unc.data_ptr(), st))
        torch.cuda.synchronize()
        bad = unc.cpu().numpy().astype(bool)
        good = todo[~bad]
        D_all[good], I_all[good] = D.cpu().numpy()[~bad], I.cpu().numpy()[~bad]
        todo = todo[bad]
        if todo.size == 0:
            break
    assert todo.size == 0, "the exact stage certifies everything"
    np.testing.assert_array_equal(I_all, Iw)
    np.testing.assert_array_equal(D_all, Dw)
    for ix in shards:
        ix.index.close()
