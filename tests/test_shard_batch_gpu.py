"""Row-sharded batches (b200_index_search_shard_dev + b200_merge_certify_dev) on ONE GPU: the G shards
are G indexes of this process searched one after the other (no kernel waits on another), their lists and bounds are
laid out as the all-gather would leave them, and the merged, certified answer is compared with the unsharded oracle —
ids and distances bit-exact.  The NCCL plumbing around the same calls is covered by tests/test_sharded_nccl_gpu.py
(>= 2 GPUs) and by the gloo test of the protocol on the CPU (tests/test_sharded_gloo_cpu.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(gpu):
    import torch

    import c99_vectordb_b200 as m
    from c99_vectordb_b200 import _cabi
    from c99_vectordb_b200.sharded import shard_range

    return m, _cabi, shard_range, torch


def sharded_batch(env, db, ids, q, k, world, metric, store="f32", normalize=False, widen=0, **opts):
    m, _cabi, shard_range, torch = env
    L = _cabi.load()
    n, d = db.shape
    nq = q.shape[0]
    dev = torch.device("cuda", 0)
    shards = []
    for g in range(world):
        lo, hi = shard_range(n, world, g)
        ix = m.IndexIDMap2(m.IndexFlat(d, metric, store=store, normalize=normalize))
        for name, v in opts.items():
            ix.index.set_option(name, v)
        if hi > lo:
            ix.add_with_ids(db[lo:hi], ids[lo:hi])
        shards.append(ix)
    qd = torch.from_numpy(q).to(dev)
    Dp = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
    Ip = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    Bp = torch.empty((world, nq), dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
    used = []
    for g, ix in enumerate(shards):
        _cabi.check(L.b200_index_search_shard_dev(ix.index._h, qd.data_ptr(), nq, k, world, widen, Dp[g].data_ptr(),
                                                  Ip[g].data_ptr(), Bp[g].data_ptr(), st))
        used.append(ix.index.get_option("stat_gemm_used"))
    D = torch.empty((nq, k), dtype=torch.float32, device=dev)
    I = torch.empty((nq, k), dtype=torch.int64, device=dev)
    unc = torch.zeros(nq, dtype=torch.int32, device=dev)
    n_unc = torch.zeros(1, dtype=torch.int32, device=dev)
    _cabi.check(L.b200_merge_certify_dev(metric, world, nq, k, n, Dp.data_ptr(), Ip.data_ptr(), 0, 0, Bp.data_ptr(), 0,
                                         D.data_ptr(), I.data_ptr(), unc.data_ptr(), n_unc.data_ptr(), st))
    torch.cuda.synchronize()
    stats = [{s: ix.index.get_option(s) for s in ("stat_gemm_pass1_us", "stat_gemm_pass2_us", "stat_gemm_rerank_us")} for ix in shards]
    for ix in shards:
        ix.index.close()
    assert int(n_unc.item()) == int(unc.sum().item())
    return D.cpu().numpy(), I.cpu().numpy(), unc.cpu().numpy().astype(bool), Bp.cpu().numpy(), used, stats


@pytest.mark.parametrize("world,metric,n,d,nq,k,store,normalize", [
    (2, 0, 140_001, 768, 130, 10, "f32", False),
    (4, 0, 400_000, 384, 256, 100, "f32", False),
    (8, 1, 330_000, 384, 64, 10, "f32", False),
    (3, 1, 200_000, 100, 97, 17, "f32", True),
    (2, 0, 150_000, 1024, 40, 10, "bf16", True),
])
def test_certified_queries_are_exact(env, world, metric, n, d, nq, k, store, normalize):
    db = oracle.synth_rows(n, d, 1234)
    db[n // 2: n // 2 + 1000] = db[:1000]  # exact ties across shards
    q = oracle.synth_rows(nq, d, 5678)
    ids = np.arange(n, dtype=np.int64) * 2 + 7
    D, I, unc, B, used, stats = sharded_batch(env, db, ids, q, k, world, metric, store, normalize, gemm_min_nq=2, gemm_min_rows=4096)
    assert all(used), used  # every shard answered on the tensor-core path
    assert all(s["stat_gemm_pass2_us"] > 0 for s in stats), stats  # the brackets of the un-synchronised form are readable
    ref_db, ref_q = db, q
    if normalize:
        ref_db = oracle.normalize_rows(db, oracle.ORDER_DEVICE)
        ref_q = oracle.normalize_rows(q, oracle.ORDER_DEVICE)
    if store == "bf16":
        ref_db = oracle.round_bf16(ref_db)
    Dw, Iw = oracle.search(metric, ref_db, ref_q, k, ids=ids, order=oracle.ORDER_DEVICE, chunk=8 if store == "bf16" else 4)
    assert unc.sum() <= nq // 4, f"{unc.sum()} of {nq} queries uncertified"
    ok = ~unc
    np.testing.assert_array_equal(I[ok], Iw[ok])
    np.testing.assert_array_equal(D[ok], Dw[ok])
    assert np.isfinite(B).all()


def test_ineligible_shards_answer_exactly_with_a_neutral_bound(env):
    """Shards too small for the tensor-core path (and one EMPTY shard) return their exact top k and a bound that
    excludes nothing: the merged answer is certified and exact."""
    n, d, nq, k, world = 2500, 64, 9, 20, 4
    db = oracle.synth_rows(n, d, 77)
    ids = np.arange(n, dtype=np.int64)
    q = oracle.synth_rows(nq, d, 78)
    for metric in (0, 1):
        D, I, unc, B, used, _ = sharded_batch(env, db, ids, q, k, world, metric)
        assert not any(used) and not unc.any()
        assert np.isinf(B).all()
        Dw, Iw = oracle.search(metric, db, q, k, ids=ids, order=oracle.ORDER_DEVICE)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    # 3 rows over 4 shards: the last shard is empty, k exceeds the database
    D, I, unc, B, used, _ = sharded_batch(env, db[:3], ids[:3], q, 5, 4, 0)
    Dw, Iw = oracle.search(0, db[:3], q, 5, ids=ids[:3], order=oracle.ORDER_DEVICE)
    assert not unc.any()
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)


def test_starved_thresholds_are_reported_not_hidden(env):
    """With the candidate budget forced far too low the lists cannot prove the top k: the certificate must say so
    (the caller retries those queries) — and whatever it does certify is still exact."""
    n, d, nq, k, world = 300_000, 256, 64, 100, 2
    db = oracle.synth_rows(n, d, 5)
    ids = np.arange(n, dtype=np.int64)
    q = oracle.synth_rows(nq, d, 6)
    D, I, unc, B, used, _ = sharded_batch(env, db, ids, q, k, world, 0, gemm_min_nq=2, gemm_emit_factor=2)
    assert all(used)
    Dw, Iw = oracle.search(0, db, q, k, ids=ids, order=oracle.ORDER_DEVICE)
    assert unc.any(), "emit_factor 2 was expected to leave queries uncertified"
    ok = ~unc
    np.testing.assert_array_equal(I[ok], Iw[ok])
    np.testing.assert_array_equal(D[ok], Dw[ok])
    # the widened second attempt certifies (most of) them
    D2, I2, unc2, *_ = sharded_batch(env, db, ids, q[unc], k, world, 0, widen=1, gemm_min_nq=2, gemm_emit_factor=2)
    assert unc2.sum() < unc.sum()
    np.testing.assert_array_equal(I2[~unc2], Iw[unc][~unc2])
