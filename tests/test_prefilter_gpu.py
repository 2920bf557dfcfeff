"""Option `prefilter` (off by default): a single query first ranks the resident bf16 shadow with the scan kernel — half
the bytes of the fp32 rows — then the best max(32, 4k) rows are re-scored from the fp32 rows and the answer is accepted
only under the certificate of DESIGN.md 7.6; otherwise the fp32 scan answers.  Either way ids and distances must be
bit-identical to the plain fp32 scan (and to the oracle's device-order restatement)."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


@pytest.mark.parametrize("n,d,k,metric,normalize,ids", [
    (300_000, 768, 10, 0, False, False),
    (200_000, 384, 10, 1, False, True),     # L2 through the augmented columns
    (150_000, 384, 1, 0, True, True),       # cosine: query normalised before the approximate scan
    (120_000, 100, 64, 1, True, False),     # kp = 256, short rows (8 lanes per row in the fp32 scan)
    (100_000, 1024, 25, 0, False, True),
])
def test_prefiltered_single_query_equals_plain_scan(b200, n, d, k, metric, normalize, ids):
    db = oracle.synth_rows(n, d, 1234)
    db[n // 2: n // 2 + 500] = db[:500]  # exact ties
    qs = oracle.synth_rows(6, d, 5678)
    idv = np.arange(n, dtype=np.int64) * 3 + 11 if ids else None
    base = b200.IndexFlat(d, metric, normalize=normalize)
    idx = base
    if ids:
        idx = b200.IndexIDMap2(base)
        idx.add_with_ids(db, idv)
    else:
        base.add(db)
    ref_db, ref_q = db, qs
    if normalize:
        ref_db = oracle.normalize_rows(db, oracle.ORDER_DEVICE)
        ref_q = oracle.normalize_rows(qs, oracle.ORDER_DEVICE)
    Dw, Iw = oracle.search(metric, ref_db, ref_q, k, ids=idv, order=oracle.ORDER_DEVICE)
    base.set_option("prefilter", 1)
    used = certified = 0
    for i in range(qs.shape[0]):
        fb0 = base.get_option("stat_prefilter_fallbacks")
        D, I = idx.search(qs[i:i + 1], k)
        used += base.get_option("stat_prefilter_used")
        certified += int(base.get_option("stat_prefilter_fallbacks") == fb0)
        np.testing.assert_array_equal(I[0], Iw[i])
        np.testing.assert_array_equal(D[0], Dw[i])
    assert used == qs.shape[0]
    assert certified >= qs.shape[0] - 1, "the certificate was expected to hold for (almost) every random query"
    base.set_option("prefilter", 0)
    D, I = idx.search(qs[:1], k)
    assert base.get_option("stat_prefilter_used") == 0
    np.testing.assert_array_equal(I[0], Iw[0])


def test_near_ties_fail_the_certificate_and_fall_back(b200):
    """Scores packed within a few ulp of one another cannot be separated by a bf16 ranking: the certificate must
    refuse and the fp32 scan must answer — identical ids."""
    n, d, k = 100_000, 128, 10
    rng = np.random.default_rng(7)
    base_row = rng.standard_normal(d).astype(np.float32)
    db = np.tile(base_row, (n, 1))
    db += (rng.standard_normal((n, d)) * 1e-6).astype(np.float32)  # 100k near-duplicates
    q = base_row[None, :].copy()
    idx = b200.IndexFlat(d, 0)
    idx.add(db)
    Dw, Iw = oracle.search(0, db, q, k, order=oracle.ORDER_DEVICE)
    idx.set_option("prefilter", 1)
    D, I = idx.search(q, k)
    assert idx.get_option("stat_prefilter_used") == 1 and idx.get_option("stat_prefilter_fallbacks") == 1
    np.testing.assert_array_equal(I, Iw)
    np.testing.assert_array_equal(D, Dw)


def test_not_applicable_cases_stay_on_the_fp32_scan(b200):
    d = 64
    db, q = oracle.synth_rows(400, d, 1), oracle.synth_rows(1, d, 2)
    idx = b200.IndexFlat(d, 0)
    idx.add(db)
    idx.set_option("prefilter", 1)
    Dw, Iw = oracle.search(0, db, q, 10, order=oracle.ORDER_DEVICE)
    D, I = idx.search(q, 10)            # too few rows for a 32-entry list to mean anything
    assert idx.get_option("stat_prefilter_used") == 0
    np.testing.assert_array_equal(I, Iw)
    mask = np.zeros(400, dtype=bool)
    mask[::3] = True
    D, I = idx.search(q, 5, row_mask=mask)  # filtered searches keep the exact scan
    assert idx.get_option("stat_prefilter_used") == 0
    D, I = idx.search(q, 100)           # 4k > 256
    assert idx.get_option("stat_prefilter_used") == 0
