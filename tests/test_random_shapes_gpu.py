"""Randomised parity sweep: ragged shapes, every metric / storage / variant / query block, with and
without ids and row masks — each case bit-exact against the oracle's device-order restatement."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

import os

CASES = int(os.environ.get("B200_RANDOM_CASES", "48"))


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


@pytest.mark.parametrize("seed", range(CASES))
def test_random_case(b200, seed):
    rng = np.random.default_rng(1000 + seed)
    d = int(rng.choice([1, 2, 3, 5, 8, 17, 31, 32, 48, 64, 96, 100, 129, 200, 260, 300, 384, 500, 768, 1000, 1024, 1536]))
    n = int(rng.integers(1, 6000)) if rng.random() < 0.7 else int(rng.integers(6000, 60000))
    nq = int(rng.choice([1, 1, 1, 2, 3, 4, 7, 8, 9, 16, 19]))
    k = int(rng.choice([1, 2, 5, 10, 33, 100, 256, 257, 1000]))
    metric = int(rng.integers(0, 2))
    store = "bf16" if rng.random() < 0.3 else "f32"
    variant = int(rng.choice([0, 1, 2]))
    use_ids = rng.random() < 0.5
    use_mask = rng.random() < 0.3
    dup = rng.random() < 0.4
    db = oracle.synth_rows(n, d, 7000 + seed)
    if dup and n > 4:
        half = n // 2
        db[half: 2 * half] = db[:half]  # exact ties everywhere
    q = oracle.synth_rows(nq, d, 9000 + seed)
    ids = (np.arange(n, dtype=np.int64) * 3 + 11) if use_ids else None
    base = b200.IndexFlat(d, metric, store=store)
    base.set_option("scan_variant", variant)
    if use_ids:
        idx = b200.IndexIDMap2(base)
        # several adds of ragged sizes, as memo's save does
        cuts = sorted(set([0, n] + rng.integers(0, n + 1, 3).tolist()))
        for a, b in zip(cuts, cuts[1:]):
            if b > a:
                idx.add_with_ids(db[a:b], ids[a:b])
    else:
        idx = base
        base.add(db)
    mask = (rng.random(n) < 0.5) if use_mask else None
    D, I = idx.search(q, k, row_mask=mask)
    ref_db = oracle.round_bf16(db) if store == "bf16" else db
    rows = np.arange(n) if mask is None else np.nonzero(mask)[0]
    ref_ids = (ids if use_ids else np.arange(n, dtype=np.int64))[rows]
    Dw, Iw = oracle.search(metric, ref_db[rows], q, k, ids=ref_ids, order=oracle.ORDER_DEVICE, chunk=8 if store == "bf16" else 4)
    desc = dict(d=d, n=n, nq=nq, k=k, metric=metric, store=store, variant=variant, ids=use_ids, mask=use_mask, dup=dup)
    np.testing.assert_array_equal(I, Iw, err_msg=str(desc))
    np.testing.assert_array_equal(D, Dw, err_msg=str(desc))


def test_add_device_tensor(b200):
    import torch

    x = torch.from_numpy(oracle.synth_rows(5000, 384, 5)).cuda()
    ids = (torch.arange(5000, dtype=torch.int64, device="cuda") * 2)
    idx = b200.IndexIDMap2(b200.IndexFlat(384, 0, normalize=True))
    idx.index.add_device(x, ids)
    want = oracle.normalize_rows(x.cpu().numpy(), oracle.ORDER_DEVICE)
    np.testing.assert_array_equal(idx.index.reconstruct_n(0, 5000), want)
    np.testing.assert_array_equal(b200.vector_to_array(idx.id_map), ids.cpu().numpy())


K3_CASES = int(os.environ.get("B200_RANDOM_K3_CASES", "16"))


@pytest.mark.parametrize("seed", range(K3_CASES))
def test_random_batched_case(b200, seed):
    """The tensor-core path over random shapes: both forms (queries resident / 256 x 256), resident and streamed shadow,
    filters, cosine, bf16 rows, id maps — ids and distances bit-exact against the oracle."""
    rng = np.random.default_rng(5000 + seed)
    d = int(rng.choice([32, 48, 64, 100, 129, 200, 256, 384, 500, 768, 1024]))
    n = int(rng.integers(8_000, 90_000))
    nq = int(rng.choice([2, 3, 5, 16, 17, 40, 100, 129, 150, 257, 300]))
    k = int(rng.choice([1, 3, 10, 16]))
    metric = int(rng.integers(0, 2))
    store = "bf16" if rng.random() < 0.25 else "f32"
    normalize = rng.random() < 0.4
    use_ids = rng.random() < 0.5
    use_mask = rng.random() < 0.3
    rows_form = int(rng.random() < 0.75)
    streamed = rng.random() < 0.3
    db = oracle.synth_rows(n, d, 17000 + seed)
    db[n // 2: n // 2 + 300] = db[:300]  # exact ties
    q = oracle.synth_rows(nq, d, 19000 + seed)
    ids = (np.arange(n, dtype=np.int64) * 2 - 5) if use_ids else None
    base = b200.IndexFlat(d, metric, store=store, normalize=normalize)
    base.set_option("gemm_rows_form", rows_form)
    if streamed:
        base.set_option("gemm_shadow_max_rows", int(rng.integers(2_000, n)))
    if use_ids:
        idx = b200.IndexIDMap2(base)
        idx.add_with_ids(db, ids)
    else:
        idx = base
        base.add(db)
    mask = (rng.random(n) < 0.6) if use_mask else None
    D, I = idx.search(q, k, row_mask=mask)
    desc = dict(d=d, n=n, nq=nq, k=k, metric=metric, store=store, normalize=normalize, ids=use_ids, mask=use_mask, rows_form=rows_form,
                streamed=streamed, used=base.get_option("stat_gemm_used"), form=base.get_option("stat_gemm_rows_form"),
                stream=base.get_option("stat_gemm_streamed"), fallbacks=base.get_option("stat_gemm_fallbacks"))
    ref_db, ref_q = db, q
    if normalize:
        ref_db = oracle.normalize_rows(db, oracle.ORDER_DEVICE)
        ref_q = oracle.normalize_rows(q, oracle.ORDER_DEVICE)
    if store == "bf16":
        ref_db = oracle.round_bf16(ref_db)
    rows = np.arange(n) if mask is None else np.nonzero(mask)[0]
    ref_ids = (ids if use_ids else np.arange(n, dtype=np.int64))[rows]
    Dw, Iw = oracle.search(metric, ref_db[rows], ref_q, k, ids=ref_ids, order=oracle.ORDER_DEVICE, chunk=8 if store == "bf16" else 4)
    np.testing.assert_array_equal(I, Iw, err_msg=str(desc))
    np.testing.assert_array_equal(D, Dw, err_msg=str(desc))
    assert desc["used"] == 1, desc
