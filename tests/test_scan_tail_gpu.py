"""The scan kernel's tail and launch plumbing (round 2): adversarial orders for the threshold-based merges,
back-to-back launches under programmatic dependent launch, fused query normalisation, stream hand-over,
negative record ids through the shard merge, phase stamps.  Everything is compared with the oracle bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


def _check(idx, metric, db, q, k, ids=None, chunk=4):
    D, I = idx.search(q, k)
    Dw, Iw = oracle.search(metric, db, q, k, ids=ids, order=oracle.ORDER_DEVICE, chunk=chunk)
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("order", ["best_last", "best_first", "all_equal", "two_values"])
def test_adversarial_row_orders(b200, variant, metric, order):
    """Worst cases for threshold-gated selection: every row better than all before it (each insert beats tau),
    the best rows packed into the first tiles (one CTA holds the whole answer), all scores equal (nothing is ever
    below a threshold: the final merge runs its multi-window path on ties), two score values."""
    n, d = 150_000, 64
    base = np.zeros((n, d), np.float32)
    ramp = np.arange(n, dtype=np.float32) / n  # exactly representable steps are not needed: the oracle sees the same rows
    if order == "best_last":
        base[:, 0] = ramp if metric == 0 else 1.0 - ramp
    elif order == "best_first":
        base[:, 0] = 1.0 - ramp if metric == 0 else ramp
    elif order == "all_equal":
        base[:, 0] = 0.5
    else:
        base[:, 0] = (np.arange(n) % 2).astype(np.float32)
    q = np.zeros((1, d), np.float32)
    q[0, 0] = 1.0 if metric == 0 else 0.0
    idx = b200.IndexFlat(d, metric)
    idx.set_option("scan_variant", variant)
    idx.add(base)
    for k in (1, 10, 100, 256):
        _check(idx, metric, base, q, k)
    idx.close()


@pytest.mark.parametrize("k", [1, 10, 200])
def test_many_back_to_back_launches_pdl(b200, k):
    """Back-to-back device-resident searches (programmatic dependent launch, both wait placements) return what
    isolated searches return: the double-buffered control words and survivor lists never mix two launches."""
    import torch

    n, d, reps = 120_000, 96, 40
    db = oracle.synth_rows(n, d, 21)
    Q = oracle.synth_rows(reps, d, 22)
    idx = b200.IndexFlat(d, 0)
    idx.add(db)
    Dw, Iw = oracle.search(0, db, Q, k, order=oracle.ORDER_DEVICE)
    qd = torch.from_numpy(Q).cuda()
    for stable in (0, 1):
        idx.set_option("queries_stable", stable)
        outs = []
        for i in range(reps):  # no synchronisation in between: launches overlap
            outs.append(idx.search_device(qd[i:i + 1], k))
        torch.cuda.synchronize()
        for i, (D, I) in enumerate(outs):
            np.testing.assert_array_equal(I.cpu().numpy()[0], Iw[i])
            np.testing.assert_array_equal(D.cpu().numpy()[0], Dw[i])
    # queries produced by the kernel right before the search (the case queries_stable = 0 exists for)
    idx.set_option("queries_stable", 0)
    src = torch.from_numpy(Q).cuda()
    outs = []
    for i in range(reps):
        qi = src[i:i + 1] * 1.0  # a torch kernel writes the query, the search follows on the same stream
        outs.append(idx.search_device(qi.contiguous(), k))
    torch.cuda.synchronize()
    for i, (D, I) in enumerate(outs):
        np.testing.assert_array_equal(I.cpu().numpy()[0], Iw[i])
    idx.close()


@pytest.mark.parametrize("d", [7, 64, 384, 1000])
@pytest.mark.parametrize("store", ["f32", "bf16"])
def test_fused_query_normalise_equals_k1(b200, d, store):
    """normalize=True indexes normalise queries inside the scan kernel's prologue; the result is bit-identical to
    normalising with K1 first (option fuse_query_normalize = 0) and to the oracle's device-order normalise."""
    n, nq = 30_000, 5
    raw = oracle.synth_rows(n, d, 31) * 3.0
    q = oracle.synth_rows(nq, d, 32) * 7.0
    q[1] = 0.0  # norm <= 1e-8 -> zero query (memo_cli.py:133)
    idx = b200.IndexFlat(d, 0, store=store, normalize=True)
    idx.add(raw)
    D1, I1 = idx.search(q[:1], 10)          # single query: fused path
    D8, I8 = idx.search(q, 10)              # query block
    idx.set_option("fuse_query_normalize", 0)
    E1, J1 = idx.search(q[:1], 10)
    E8, J8 = idx.search(q, 10)
    np.testing.assert_array_equal(I1, J1)
    np.testing.assert_array_equal(D1, E1)
    np.testing.assert_array_equal(I8, J8)
    np.testing.assert_array_equal(D8, E8)
    db = oracle.normalize_rows(raw, oracle.ORDER_DEVICE)
    if store == "bf16":
        db = oracle.round_bf16(db)
    qn = oracle.normalize_rows(q, oracle.ORDER_DEVICE)
    Dw, Iw = oracle.search(0, db, qn, 10, order=oracle.ORDER_DEVICE, chunk=8 if store == "bf16" else 4)
    np.testing.assert_array_equal(I8, Iw)
    np.testing.assert_array_equal(D8, Dw)
    idx.close()


def test_stream_hand_over(b200):
    """Searches enqueued on different streams share the handle's scratch: the library orders each new stream behind
    the previous one, so alternating streams (and the host API in between) never corrupt a result."""
    import torch

    n, d, k = 400_000, 128, 10
    db = oracle.synth_rows(n, d, 41)
    Q = oracle.synth_rows(12, d, 42)
    Dw, Iw = oracle.search(0, db, Q, k, order=oracle.ORDER_DEVICE)
    idx = b200.IndexFlat(d, 0)
    idx.add(db)
    qd = torch.from_numpy(Q).cuda()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(12):
        st = (s1, s2)[i % 2]
        with torch.cuda.stream(st):
            outs.append(idx.search_device(qd[i:i + 1], k))
        if i == 5:
            Dh, Ih = idx.search(Q[3:4], k)  # host entry on the handle's own stream in the middle
            np.testing.assert_array_equal(Ih[0], Iw[3])
    torch.cuda.synchronize()
    for i, (D, I) in enumerate(outs):
        np.testing.assert_array_equal(I.cpu().numpy()[0], Iw[i])
        np.testing.assert_array_equal(D.cpu().numpy()[0], Dw[i])
    idx.close()


@pytest.mark.parametrize("metric", [0, 1])
def test_merge_keeps_negative_record_ids(b200, metric):
    """K4 recognises padding by its sentinel score, not by id < 0: negative record ids (legal in an id map) survive
    the shard merge, while real padding still sorts last."""
    import torch

    from c99_vectordb_b200 import _cabi

    rng = np.random.default_rng(5)
    G, nq, k = 3, 4, 6
    D = np.sort(rng.standard_normal((G, nq, k)).astype(np.float32), axis=2)
    if metric == 0:
        D = D[:, :, ::-1].copy()
    I = -rng.integers(1, 1000, size=(G, nq, k)).astype(np.int64)  # every id negative
    pad = -np.finfo(np.float32).max if metric == 0 else np.finfo(np.float32).max
    D[1, :, 4:] = pad  # shard 1 ran out of candidates
    I[1, :, 4:] = -1
    Dw, Iw = oracle.merge_topk(metric, D, I)
    Dd, Id = torch.from_numpy(D).cuda(), torch.from_numpy(I).cuda()
    Do = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    Io = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    _cabi.check(_cabi.load().b200_merge_topk_dev(metric, G, nq, k, Dd.data_ptr(), Id.data_ptr(), 0, 0, Do.data_ptr(), Io.data_ptr(),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(Io.cpu().numpy(), Iw)
    np.testing.assert_array_equal(Do.cpu().numpy(), Dw)
    assert (Iw < 0).all() and (Dw != pad).all()  # k = 6 <= 6 + 4 + 6 real candidates: no padding in the merged lists


def test_negative_ids_through_index(b200):
    n, d = 5000, 32
    db = oracle.synth_rows(n, d, 51)
    ids = -(np.arange(n, dtype=np.int64) * 7 + 3)
    idx = b200.IndexIDMap2(b200.IndexFlat(d, 1))
    idx.add_with_ids(db, ids)
    _check(idx, 1, db, oracle.synth_rows(3, d, 52), 10, ids=ids)


def test_phase_stamps(b200):
    """Option scan_phase_stamps: every CTA stamps its phases in order; exactly one CTA runs the final merge."""
    import torch

    from c99_vectordb_b200 import _cabi

    idx = b200.IndexFlat(256, 0)
    idx.add_synthetic(500_000, 1234)
    idx.set_option("scan_phase_stamps", 1)
    q = torch.from_numpy(oracle.synth_rows(1, 256, 1)).cuda()
    idx.search_device(q, 10)
    torch.cuda.synchronize()
    buf = np.zeros(1024 * 8, dtype=np.uint64)
    n_ctas = C.c_int64(0)
    _cabi.check(_cabi.load().b200_index_read_phase_stamps(idx._h, buf.ctypes.data, buf.size, C.byref(n_ctas)))
    st = buf[: n_ctas.value * 8].reshape(-1, 8).astype(np.int64)
    assert n_ctas.value >= 100
    assert (st[:, 0] > 0).all() and (st[:, 1] >= st[:, 0]).all() and (st[:, 2] >= st[:, 1]).all() and (st[:, 3] >= st[:, 2]).all()
    last = st[:, 6] > 0
    assert last.sum() == 1 and (st[last, 6] >= st[last, 5]).all() and (st[last, 5] >= st[last, 4]).all()
    idx.set_option("scan_phase_stamps", 0)
    idx.close()
