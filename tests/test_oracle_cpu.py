"""CPU tests of the oracle itself: against the reference-generated golden vectors, against an
independent numpy fp64 twin, and on the adversarial cases of SURVEY.md §4."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle

FLT_MAX = np.finfo(np.float32).max


def test_normalize_matches_reference_golden():
    g = np.load(GOLDEN / "normalize.npz")
    for i in range(int(g["n"])):
        vin, ref = g[f"in_{i}"], g[f"out_{i}"]
        for order in (oracle.ORDER_SIMD, oracle.ORDER_DEVICE):
            out = oracle.normalize_rows(vin, order)
            # summation order differs from numpy's BLAS sdot: 2 ulp of slack on unit-scale values
            np.testing.assert_allclose(out, ref, rtol=3e-7, atol=1e-12, err_msg=f"case {i} order {order}")
        if not ref.any():
            assert not oracle.normalize_rows(vin).any()  # zero-vector rule must hold exactly


def test_normalize_threshold_is_a_double_compare():
    # float32(1e-8) < 1e-8 (double) -> zeros; the next float up is > 1e-8 -> kept (memo_cli.py:133)
    v = np.array([np.float32(1e-8), 0, 0, 0], np.float32)
    assert not oracle.normalize_rows(v).any()
    w = np.array([np.nextafter(np.float32(1e-8), np.float32(1)), 0, 0, 0], np.float32)
    assert oracle.normalize_rows(w)[0] == 1.0


def test_synth_is_exact_and_in_range():
    a = oracle.synth_rows(1000, 37, seed=1234)
    b = oracle.synth_rows(500, 37, seed=1234, first_row=500)
    assert a.dtype == np.float32 and a.min() >= -1.0 and a.max() < 1.0
    np.testing.assert_array_equal(a[500:], b)  # counter based: any window is reproducible
    assert abs(float(a.mean())) < 0.02 and abs(float(a.std()) - 3 ** -0.5) < 0.02
    assert np.all(a * 8388608.0 == np.round(a * 8388608.0))  # multiples of 2^-23


def test_round_bf16_matches_torch():
    import torch

    x = np.concatenate([oracle.synth_rows(64, 33, 7).ravel(), np.array([0.0, -0.0, 1.0, 1.00390625, 1.005859375, 3.3895314e38, 1e-40], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(oracle.round_bf16(x), want)


@pytest.mark.parametrize("metric", [oracle.METRIC_IP, oracle.METRIC_L2])
@pytest.mark.parametrize("n,d,k,nq", [(1, 1, 1, 1), (17, 3, 5, 2), (1000, 384, 10, 7), (300, 768, 100, 3), (50, 100, 80, 4)])
def test_search_against_fp64_twin(metric, n, d, k, nq):
    db = oracle.synth_rows(n, d, 11)
    q = oracle.synth_rows(nq, d, 22)
    for order, chunk in ((oracle.ORDER_SIMD, 4), (oracle.ORDER_DEVICE, 4), (oracle.ORDER_DEVICE, 8)):
        D, I = oracle.search(metric, db, q, k, order=order, chunk=chunk)
        for i in range(nq):
            assert oracle.check_topk_against_truth(metric, db, q[i], D[i], I[i]) == []
    D64, I64 = oracle.np_search_f64(metric, db, q, k)
    D, I = oracle.search(metric, db, q, k)
    m = min(k, n)
    # random data has no near-ties at these sizes: ids must agree exactly with the fp64 ranking
    np.testing.assert_array_equal(I[:, :m], I64[:, :m])
    np.testing.assert_allclose(D[:, :m], D64[:, :m], rtol=1e-5, atol=1e-5)


def test_rowpar_equals_sequential():
    db, q = oracle.synth_rows(5000, 96, 5), oracle.synth_rows(3, 96, 6)
    for metric in (0, 1):
        a = oracle.search(metric, db, q, 10)
        b = oracle.search(metric, db, q, 10, rowpar=True)
        np.testing.assert_array_equal(a[1], b[1])
        np.testing.assert_array_equal(a[0], b[0])


@pytest.mark.parametrize("metric", [0, 1])
def test_exact_ties_smaller_row_first(metric):
    base = oracle.synth_rows(8, 16, 3)
    db = np.concatenate([base, base, base[:3]], axis=0)  # rows i, i+8 (and i+16 for i<3) identical
    q = oracle.synth_rows(1, 16, 4)
    D, I = oracle.search(metric, db, q, 19)
    for a, b in zip(range(18), range(1, 19)):
        if D[0, a] == D[0, b]:
            assert I[0, a] < I[0, b]
    # boundary: k cuts through a tie group -> the smaller rows are kept
    best = I[0, 0] % 8
    D2, I2 = oracle.search(metric, db, q, 1)
    assert I2[0, 0] == best


@pytest.mark.parametrize("metric", [0, 1])
def test_padding_and_k_larger_than_ntotal(metric):
    db, q = oracle.synth_rows(3, 8, 1), oracle.synth_rows(2, 8, 2)
    D, I = oracle.search(metric, db, q, 6)
    assert (I[:, 3:] == -1).all() and (I[:, :3] >= 0).all()
    assert (D[:, 3:] == (-FLT_MAX if metric == 0 else FLT_MAX)).all()
    D0, I0 = oracle.search(metric, np.zeros((0, 8), np.float32), q, 4)
    assert (I0 == -1).all()


@pytest.mark.parametrize("metric", [0, 1])
def test_nan_and_inf_rows_never_enter(metric):
    db = oracle.synth_rows(10, 8, 1)
    db[2, 3] = np.nan
    db[5, 0] = np.inf if metric == 1 else -np.inf
    q = np.abs(oracle.synth_rows(1, 8, 2)) + 0.1
    D, I = oracle.search(metric, db, q, 10)
    got = set(I[0][I[0] >= 0].tolist())
    assert 2 not in got and 5 not in got and len(got) == 8 and (I[0, 8:] == -1).all()


def test_zero_vectors_and_idmap():
    db = np.zeros((4, 8), np.float32)
    db[2] = 1.0
    ids = np.array([100, 7, 42, 9], np.int64)
    q = np.ones((1, 8), np.float32)
    D, I = oracle.search(0, db, q, 4, ids=ids)
    assert I[0].tolist() == [42, 100, 7, 9] and D[0].tolist() == [8.0, 0.0, 0.0, 0.0]
    D, I = oracle.search(1, db, q, 2, ids=ids)
    assert I[0].tolist() == [42, 100] and D[0].tolist() == [0.0, 8.0]


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_sharded_merge_equals_unsharded(metric, G):
    n, d, k, nq = 1003, 32, 10, 4
    base = oracle.synth_rows(n // 2, d, 9)
    db = np.concatenate([base, base, oracle.synth_rows(n - 2 * (n // 2), d, 10)])  # cross-shard ties
    q = oracle.synth_rows(nq, d, 8)
    want = oracle.search(metric, db, q, k)
    per = -(-n // G)
    Dp, Ip = [], []
    for g in range(G):
        lo, hi = g * per, min(n, (g + 1) * per)
        Dg, Ig = oracle.search(metric, db[lo:hi], q, k, ids=np.arange(lo, hi, dtype=np.int64))
        Dp.append(Dg), Ip.append(Ig)
    Dm, Im = oracle.merge_topk(metric, np.stack(Dp), np.stack(Ip))
    np.testing.assert_array_equal(Im, want[1])
    np.testing.assert_array_equal(Dm, want[0])


def test_adapter_golden_is_self_consistent():
    """The reference's search_all over the oracle-backed stub == the oracle called directly."""
    g = np.load(GOLDEN / "adapter.npz", allow_pickle=True)
    kept, vecs = g["kept"], g["kept_vectors"]
    np.testing.assert_array_equal(g["existing_ids"], kept)
    D, I = oracle.search(oracle.METRIC_L2, vecs, g["qvecs"], len(kept), ids=kept)
    np.testing.assert_array_equal(I, g["res_ids"])
    np.testing.assert_array_equal(D, g["res_scores"])


@pytest.mark.parametrize("metric", [oracle.METRIC_IP, oracle.METRIC_L2])
def test_oracle_against_independent_sklearn_bruteforce(metric):
    """faiss is not installable here, so as a second independent implementation the oracle's ids are
    checked against scikit-learn's exact brute-force k-NN (float64) on well separated random data."""
    sk = pytest.importorskip("sklearn.neighbors")
    n, d, k, nq = 5000, 64, 10, 20
    db = oracle.normalize_rows(oracle.synth_rows(n, d, 41))
    q = oracle.normalize_rows(oracle.synth_rows(nq, d, 42))
    D, I = oracle.search(metric, db, q, k)
    # on unit vectors L2-ascending and IP-descending rankings coincide
    nn = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(db.astype(np.float64))
    dist, ind = nn.kneighbors(q.astype(np.float64))
    np.testing.assert_array_equal(I, ind)
    if metric == oracle.METRIC_L2:
        np.testing.assert_allclose(D, dist ** 2, rtol=1e-5, atol=1e-6)
    else:
        np.testing.assert_allclose(D, 1.0 - dist ** 2 / 2.0, rtol=1e-5, atol=1e-6)
