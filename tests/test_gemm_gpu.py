"""K3 (tcgen05 batched path) parity: ids and distances must be IDENTICAL to the exact scan path's
contract — bit-exact against the oracle's device-order restatement — because every emitted
candidate is re-scored in fp32 and uncertified queries are recomputed by the scan kernel."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


def run_case(m, n, d, nq, k, store="f32", normalize=False, dup=False, ids=False, metric=0, **opts):
    db = oracle.synth_rows(n, d, 1234)
    if dup:
        db[n // 2: n // 2 + 1000] = db[:1000]  # exact ties across tiles
    q = oracle.synth_rows(nq, d, 5678)
    if normalize:
        db = oracle.normalize_rows(db, oracle.ORDER_DEVICE)
        q = oracle.normalize_rows(q, oracle.ORDER_DEVICE)
    base = m.IndexFlat(d, metric, store=store)
    base.set_option("gemm_min_nq", 32)
    for name, v in opts.items():
        base.set_option(name, v)
    idv = None
    if ids:
        idv = np.arange(n, dtype=np.int64) * 2 + 7
        w = m.IndexIDMap2(base)
        w.add_with_ids(db, idv)
        idx = w
    else:
        base.add(db)
        idx = base
    D, I = idx.search(q, k)
    ref_db = oracle.round_bf16(db) if store == "bf16" else db
    Dw, Iw = oracle.search(metric, ref_db, q, k, ids=idv, order=oracle.ORDER_DEVICE, chunk=8 if store == "bf16" else 4)
    stats = {s: base.get_option(s) for s in ("stat_gemm_used", "stat_gemm_fallbacks", "stat_gemm_cand_total",
                                             "stat_gemm_pass1_us", "stat_gemm_pass2_us", "stat_gemm_rerank_us")}
    np.testing.assert_array_equal(I, Iw, err_msg=str(stats))
    np.testing.assert_array_equal(D, Dw, err_msg=str(stats))
    return stats


@pytest.mark.parametrize("n,d,nq,k", [(70001, 768, 130, 10), (200_000, 768, 256, 100), (131072, 384, 64, 10),
                                        (100_000, 100, 97, 17), (300_000, 1024, 33, 256)])
def test_batched_ip_exact(b200, n, d, nq, k):
    st = run_case(b200, n, d, nq, k)
    assert st["stat_gemm_used"] == 1
    assert st["stat_gemm_fallbacks"] <= nq // 4, st  # the certificate passes for almost every query


def test_too_few_rows_for_k_stays_on_the_scan_path(b200):
    st = run_case(b200, 66000, 1024, 33, 256)  # n < 512 k: the threshold statistic cannot resolve k
    assert st["stat_gemm_used"] == 0


def test_batched_normalized_with_ties_and_ids(b200):
    st = run_case(b200, 150_000, 768, 200, 10, normalize=True, dup=True, ids=True)
    assert st["stat_gemm_used"] == 1


def test_batched_bf16_store(b200):
    st = run_case(b200, 120_000, 1024, 96, 10, store="bf16", normalize=True)
    assert st["stat_gemm_used"] == 1


def test_uncertified_queries_fall_back_to_exact_scan(b200):
    # an emission target of ~2k candidates cannot be certified against the bf16 bound: every query
    # must be recomputed by the scan kernel and the answer must still be exact
    st = run_case(b200, 80_000, 256, 40, 50, gemm_emit_factor=2, gemm_sample_tiles=64)
    assert st["stat_gemm_fallbacks"] > 0, st
    assert st["stat_gemm_used"] == 1


def test_small_batches_do_not_use_gemm(b200):
    db, q = oracle.synth_rows(70_000, 128, 1), oracle.synth_rows(8, 128, 2)
    idx = b200.IndexFlat(128, 0)
    idx.add(db)
    idx.search(q[:1], 5)
    assert idx.get_option("stat_gemm_used") == 0  # single queries stay on the fp32 scan (gemm_min_nq = 2)
    D, I = idx.search(q, 5)
    assert idx.get_option("stat_gemm_used") == 1  # measured crossover: 8 queries already favour the tensor cores 5x
    Dw, Iw = oracle.search(0, db, q, 5, order=oracle.ORDER_DEVICE)
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)


@pytest.mark.parametrize("n,d,nq,k,normalize", [(120_000, 384, 100, 10, True), (200_000, 768, 130, 100, False),
                                                  (90_000, 100, 64, 5, False), (150_000, 62, 70, 10, True)])
def test_batched_l2_exact(b200, n, d, nq, k, normalize):
    """memo's default metric (squared L2): the GEMM ranks q.y - |y|^2/2 through two extra K columns."""
    st = run_case(b200, n, d, nq, k, metric=1, normalize=normalize, dup=True, ids=True)
    assert st["stat_gemm_used"] == 1
    assert st["stat_gemm_fallbacks"] <= nq // 4, st


def test_batched_l2_bf16_store(b200):
    st = run_case(b200, 100_000, 512, 80, 10, metric=1, store="bf16")
    assert st["stat_gemm_used"] == 1


@pytest.mark.parametrize("n,d,nq,k,metric", [(10_000, 384, 100, 10, 0), (5_200, 384, 37, 10, 1), (4_096, 64, 40, 5, 0), (20_000, 768, 50, 30, 1)])
def test_small_databases_use_the_tensor_path_too(b200, n, d, nq, k, metric):
    """BASELINE config 0 shape (10k x 384, 100 queries, k=10): one K3 sweep instead of 13 scan passes."""
    st = run_case(b200, n, d, nq, k, metric=metric, normalize=True)
    assert st["stat_gemm_used"] == 1
    assert st["stat_gemm_fallbacks"] <= max(2, nq // 4), st


@pytest.mark.parametrize("metric", [0, 1])
def test_batched_path_with_nan_inf_and_huge_rows(b200, metric):
    """NaN rows never enter, +-inf follow the scan path's rules, and magnitudes that overflow bf16 make
    the certificate fail safely (recomputed exactly) — results must equal the scan path bit for bit."""
    n, d, nq, k = 70_000, 64, 9, 10
    db = oracle.synth_rows(n, d, 77)
    db[5, 3] = np.nan
    db[6, 0] = np.inf
    db[7, 0] = -np.inf
    db[8] *= 1e30          # large but finite scores
    db[9, 1] = 3.0e38      # rounds to inf in bf16
    db[n - 1, 5] = np.nan
    q = np.abs(oracle.synth_rows(nq, d, 78)) + 0.05
    idx = b200.IndexFlat(d, metric)
    idx.add(db)
    D, I = idx.search(q, k)
    assert idx.get_option("stat_gemm_used") == 1
    idx.set_option("gemm_min_nq", 0)  # the scan path as the reference behaviour
    Ds, Is = idx.search(q, k)
    assert idx.get_option("stat_gemm_used") == 0
    np.testing.assert_array_equal(I, Is)
    np.testing.assert_array_equal(D, Ds)
    Dw, Iw = oracle.search(metric, db, q, k, order=oracle.ORDER_DEVICE)
    np.testing.assert_array_equal(I, Iw)
    assert not np.isin(I, [5, n - 1]).any()


def test_shadow_follows_adds_and_reset(b200):
    d = 128
    a, b = oracle.synth_rows(50_000, d, 1), oracle.synth_rows(30_000, d, 2)
    q = oracle.synth_rows(6, d, 3)
    idx = b200.IndexFlat(d, 0)
    idx.add(a)
    D1, I1 = idx.search(q, 10)
    assert idx.get_option("stat_gemm_used") == 1
    idx.add(b)  # the bf16 shadow must be rebuilt to cover the new rows
    D2, I2 = idx.search(q, 10)
    Dw, Iw = oracle.search(0, np.concatenate([a, b]), q, 10, order=oracle.ORDER_DEVICE)
    np.testing.assert_array_equal(I2, Iw)
    np.testing.assert_array_equal(D2, Dw)
    idx.reset()
    idx.add(b)
    D3, I3 = idx.search(q, 10)
    Dw3, Iw3 = oracle.search(0, b, q, 10, order=oracle.ORDER_DEVICE)
    np.testing.assert_array_equal(I3, Iw3)


@pytest.mark.parametrize("metric", [0, 1])
def test_batched_masked_search(b200, metric):
    """Filter push-down on the tensor-core path: excluded rows never become candidates."""
    n, d, nq, k = 90_000, 128, 40, 10
    db, q = oracle.synth_rows(n, d, 5), oracle.synth_rows(nq, d, 6)
    db[1000:1100] = db[2000:2100]
    ids = np.arange(n, dtype=np.int64) + 7
    idx = b200.IndexIDMap2(b200.IndexFlat(d, metric))
    idx.add_with_ids(db, ids)
    rng = np.random.default_rng(9)
    for frac in (0.5, 0.05):
        mask = rng.random(n) < frac
        rows = np.nonzero(mask)[0]
        D, I = idx.search(q, k, row_mask=mask)
        assert idx.index.get_option("stat_gemm_used") == 1
        Dw, Iw = oracle.search(metric, db[rows], q, k, ids=ids[rows], order=oracle.ORDER_DEVICE)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    # fewer allowed rows than k: the certificate cannot hold, the exact scan pads with -1
    mask = np.zeros(n, dtype=bool)
    mask[[5, 77, 4000]] = True
    D, I = idx.search(q, k, row_mask=mask)
    assert (I[:, 3:] == -1).all() and set(I[0, :3].tolist()) == {12, 84, 4007}


@pytest.mark.parametrize("n,d,nq,k,metric,store,cap", [
    (200_000, 384, 100, 10, 1, "f32", 65_536),   # 4 chunks, the last one ragged (3392 rows = 13.25 tiles)
    (150_001, 768, 64, 100, 0, "f32", 32_768),   # odd row count: a ragged tile inside the last chunk
    (120_000, 100, 70, 10, 0, "f32", 20_000),    # cap rounds down to whole 256-row tiles; vector path with zero padded K
    (90_000, 50, 40, 5, 1, "f32", 30_000),       # d % 4 != 0: the scalar conversion path
    (100_000, 512, 48, 10, 0, "bf16", 25_600),   # bf16 rows
])
def test_streamed_shadow_when_it_does_not_fit(b200, n, d, nq, k, metric, store, cap):
    """When the bf16 shadow cannot stay resident (100M x 384 fp32 fills the GPU; here: option gemm_shadow_max_rows) the
    rows are rounded chunk by chunk into an L2-sized scratch and swept while hot — same candidates, same exact re-rank,
    same certificate, so ids and distances stay bit-identical to the scan path."""
    st = run_case(b200, n, d, nq, k, metric=metric, store=store, dup=True, ids=True, gemm_shadow_max_rows=cap)
    assert st["stat_gemm_used"] == 1
    assert st["stat_gemm_fallbacks"] <= nq // 4, st


def test_streamed_shadow_masked_and_toggled(b200):
    n, d, nq, k = 130_000, 128, 33, 10
    db, q = oracle.synth_rows(n, d, 15), oracle.synth_rows(nq, d, 16)
    idx = b200.IndexFlat(d, 0)
    idx.add(db)
    Dw, Iw = oracle.search(0, db, q, k, order=oracle.ORDER_DEVICE)
    for cap in (0, 40_000, 0, 16_384):  # resident -> streamed -> resident -> streamed with another chunk size
        idx.set_option("gemm_shadow_max_rows", cap)
        D, I = idx.search(q, k)
        assert idx.get_option("stat_gemm_used") == 1 and idx.get_option("stat_gemm_streamed") == (1 if cap else 0)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    mask = np.random.default_rng(3).random(n) < 0.3
    rows = np.nonzero(mask)[0]
    D, I = idx.search(q, k, row_mask=mask)
    assert idx.get_option("stat_gemm_streamed") == 1
    Dm, Im = oracle.search(0, db[rows], q, k, ids=rows.astype(np.int64), order=oracle.ORDER_DEVICE)
    np.testing.assert_array_equal(I, Im)
    np.testing.assert_array_equal(D, Dm)
