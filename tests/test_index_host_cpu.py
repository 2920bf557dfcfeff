"""Host logic of c99_vectordb_b200/index.py and memo_adapter.py on CPU: the index entry points of the
C ABI are replaced by tests/fake_cabi.py (test infrastructure, numpy brute force), so what runs here
is the Python side — `.memo` headers (faiss layout), id maps, memo's adapter
functions against the goldens recorded from the unmodified reference, error conversion.  The same
scenarios run bit-exact against the CUDA library in the -m gpu tests."""
import struct
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_cabi  # noqa: E402

from c99_vectordb_b200 import _cabi  # noqa: E402
from c99_vectordb_b200 import index as ix  # noqa: E402
from c99_vectordb_b200 import memo_adapter as ma  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(autouse=True)
def fake(monkeypatch):
    real = _cabi.load()
    monkeypatch.setattr(_cabi, "_lib", fake_cabi.FakeLib(real))
    yield
    # monkeypatch restores the real library


def _hdr(d, n, metric):
    return struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric)


def _rows(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def test_write_index_is_faiss_layout_and_reads_back(tmp_path):
    d, n = 12, 37
    x, ids = _rows(n, d, 1), np.arange(n, dtype=np.int64) * 7 + 3
    idx = ix.IndexIDMap2(ix.IndexFlatL2(d))
    idx.add_with_ids(x[:20], ids[:20])
    idx.add_with_ids(x[20:], ids[20:])
    p = tmp_path / "a.memo"
    ix.write_index(idx, str(p))
    want = (b"IxM2" + _hdr(d, n, 1) + b"IxF2" + _hdr(d, n, 1) + struct.pack("<Q", n * d) + x.tobytes()
            + struct.pack("<Q", n) + ids.tobytes())
    assert p.read_bytes() == want
    back = ix.read_index(str(p))
    assert isinstance(back, ix.IndexIDMap2) and back.ntotal == n and back.d == d and back.metric_type == ix.METRIC_L2
    np.testing.assert_array_equal(ix.vector_to_array(back.id_map), ids)
    np.testing.assert_array_equal(back.index.reconstruct_n(0, n), x)
    np.testing.assert_array_equal(back.reconstruct(int(ids[5])), x[5])
    with pytest.raises(RuntimeError, match="not found"):
        back.reconstruct(10**9)
    flat = ix.IndexFlatIP(d)
    flat.add(x)
    ix.write_index(flat, str(p))
    assert p.read_bytes() == b"IxFI" + _hdr(d, n, 0) + struct.pack("<Q", n * d) + x.tobytes()
    back = ix.read_index(str(p))
    assert type(back) is ix.IndexFlat and back.metric_type == ix.METRIC_INNER_PRODUCT and back.ntotal == n
    empty = ix.IndexIDMap2(ix.IndexHNSWFlat(384, 32))
    ix.write_index(empty, str(p))
    assert ix.read_index(str(p)).ntotal == 0


def test_memo_hnsw_file_is_imported_without_the_graph(tmp_path):
    d, n = 8, 5
    x, ids = _rows(n, d, 2), np.array([0, 2, 4, 6, 9], dtype=np.int64)

    def vec(a):
        return struct.pack("<Q", len(a)) + a.tobytes()

    blob = b"IxM2" + _hdr(d, n, 1) + b"IHNf" + _hdr(d, n, 1)
    blob += vec(np.array([0.9, 0.1])) + vec(np.array([0, 64, 96], dtype=np.int32)) + vec(np.ones(n, dtype=np.int32))
    blob += vec(np.arange(n + 1, dtype=np.uint64) * 64) + vec(np.full(n * 64, -1, dtype=np.int32)) + struct.pack("<5i", 0, 0, 200, 64, 1)
    blob += b"IxF2" + _hdr(d, n, 1) + struct.pack("<Q", n * d) + x.tobytes() + struct.pack("<Q", n) + ids.tobytes()
    p = tmp_path / "legacy.memo"
    p.write_bytes(blob)
    idx = ix.read_index(str(p))
    assert isinstance(idx, ix.IndexIDMap2) and idx.ntotal == n
    np.testing.assert_array_equal(ix.vector_to_array(idx.id_map), ids)
    np.testing.assert_array_equal(idx.index.reconstruct_n(0, n), x)
    # levels vector that disagrees with ntotal -> not importable
    bad = blob.replace(vec(np.ones(n, dtype=np.int32)), vec(np.ones(n + 1, dtype=np.int32)), 1)
    p.write_bytes(bad)
    with pytest.raises(RuntimeError, match="levels"):
        ix.read_index(str(p))


@pytest.mark.parametrize("damage", ["fourcc", "count", "short_rows", "short_ids", "id_count", "header", "empty"])
def test_unreadable_files_raise_and_memo_starts_fresh(tmp_path, damage):
    """read_index raises; load_index (memo_cli.py:251-261) then returns a fresh empty index."""
    d, n = 6, 9
    x, ids = _rows(n, d, 3), np.arange(n, dtype=np.int64)
    good = (b"IxM2" + _hdr(d, n, 1) + b"IxF2" + _hdr(d, n, 1) + struct.pack("<Q", n * d) + x.tobytes()
            + struct.pack("<Q", n) + ids.tobytes())
    flat_at = 4 + 33
    blob = {
        "fourcc": b"IwFl" + good[4:],
        "count": good[:flat_at + 4 + 33] + struct.pack("<Q", n * d + 1) + good[flat_at + 4 + 33 + 8:],
        "short_rows": good[: flat_at + 4 + 33 + 8 + 10],
        "short_ids": good[:-8],
        "id_count": good[: -(8 + n * 8)] + struct.pack("<Q", n - 1) + ids.tobytes(),
        "header": good[:20],
        "empty": b"",
    }[damage]
    p = tmp_path / "bad.memo"
    p.write_bytes(blob)
    with pytest.raises(RuntimeError):
        ix.read_index(str(p))
    fresh = ma.load_index(p)
    assert isinstance(fresh, ix.IndexIDMap2) and fresh.ntotal == 0
    with pytest.raises(FileNotFoundError):
        ix.read_index(str(tmp_path / "missing.memo"))
    assert ma.load_index(tmp_path / "missing.memo").ntotal == 0


def test_faiss_like_argument_errors():
    idx = ix.IndexIDMap2(ix.IndexFlatL2(4))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 5), np.float32), np.arange(2))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 4), np.float32), np.arange(3))
    with pytest.raises(RuntimeError, match="add_with_ids"):
        idx.add(np.zeros((1, 4), np.float32))
    with pytest.raises(RuntimeError, match="add_with_ids not implemented"):
        ix.IndexFlatL2(4).add_with_ids(np.zeros((1, 4), np.float32), np.arange(1))
    full = ix.IndexFlatL2(4)
    full.add(np.zeros((1, 4), np.float32))
    with pytest.raises(RuntimeError, match="empty"):
        ix.IndexIDMap2(full)
    with pytest.raises(ValueError):
        ix.IndexFlat(4, store="fp8")
    with pytest.raises(RuntimeError, match="positive"):  # C-ABI status -> RuntimeError with the library's text
        ix.IndexFlat(0)
    with pytest.raises(RuntimeError, match="serialize"):
        ix.write_index(object(), "/tmp/never-written.memo")


def test_row_mask_packing_and_filters():
    m = np.zeros(70, dtype=bool)
    m[[0, 31, 32, 69]] = True
    w = ix.pack_row_mask(m, 70)
    assert w.dtype == np.dtype("<u4") and w.tolist() == [(1 << 0) | (1 << 31), 1, 1 << 5]
    with pytest.raises(AssertionError):
        ix.pack_row_mask(m, 71)
    d, n = 5, 70
    x, ids = _rows(n, d, 4), np.arange(n, dtype=np.int64) + 100
    idx = ix.IndexIDMap2(ix.IndexFlatL2(d))
    idx.add_with_ids(x, ids)
    D, I = idx.search(x[:2], 3, row_mask=m)
    assert set(I.ravel().tolist()) <= {100, 131, 132, 169}
    D, I = idx.search(x[:2], 3, ids_allowed=[100, 101, 5])
    assert set(I.ravel().tolist()) == {100, 101, -1}
    with pytest.raises(ValueError):
        idx.search(x[:1], 1, row_mask=m, ids_allowed=[1])
    D, I = idx.search(x[:1], n + 4)
    assert D.shape == (1, n + 4) and (I[0, n:] == -1).all() and I[0, 0] == 100


def test_memo_adapter_against_reference_golden_on_cpu():
    """rebuild_index_from_texts / get_existing_ids / search_all (memo_cli.py:265-298) against what the
    unmodified reference returned (tests/golden/make_golden.py); ranking compared modulo near ties."""
    g = np.load(GOLDEN / "adapter.npz", allow_pickle=True)
    records = [None if r == "\x00NONE" else r for r in g["records"].tolist()]
    idx = ma.rebuild_index_from_texts(records, vectors=g["kept_vectors"])
    assert ma.get_existing_ids(idx) == set(g["kept"].tolist())
    for qv, ids_ref, sc_ref in zip(g["qvecs"], g["res_ids"], g["res_scores"]):
        res = ma.search_all(idx, qv)
        assert len(res) == len(ids_ref)
        np.testing.assert_allclose([r.score for r in res], sc_ref, rtol=1e-5, atol=1e-6)
        got, start = [r.doc_id for r in res], 0
        for i in range(1, len(ids_ref) + 1):
            if i == len(ids_ref) or abs(float(sc_ref[i]) - float(sc_ref[i - 1])) > 2e-6:
                assert sorted(got[start:i]) == sorted(np.asarray(ids_ref[start:i]).tolist())
                start = i


def test_add_texts_host_side_packs_and_skips_blanks():
    """pack_texts / IndexIDMap2.add_texts (the host half of K6): ASCII corpora travel as one blob with the blank
    detection left to the engine, non-ASCII corpora are lower-cased and filtered per record; either way the rows and
    ids equal the reference-shaped rebuild with the stable hash injected."""
    ascii_texts = ["Peanuts ALLERGY note", "", "   \t\n", "wifi password hunter2", None, "!!! ...", "snake_case x1 X1"]
    uni_texts = ["café Kelvin K", "  ", "Straße ÄÖÜ", "", "plain ascii too"]
    for texts in (ascii_texts, uni_texts):
        idx = ma.create_index()
        added = idx.add_texts(texts)
        want_keep = [i for i, t in enumerate(texts) if not ma.is_blank_body(t or "")]
        assert added == len(want_keep) == idx.ntotal
        np.testing.assert_array_equal(ix.vector_to_array(idx.id_map), np.asarray(want_keep, dtype=np.int64))
        raw = ma.embed_texts([texts[i] or "" for i in want_keep], hash_fn=ma.stable_hash)
        nrm = np.sqrt((raw.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
        want = np.where(nrm <= 1e-8, 0, raw / np.maximum(nrm, 1e-30)).astype(np.float32)
        np.testing.assert_array_equal(idx.index.reconstruct_n(0, idx.ntotal), want)
    blob, offsets, keep_ids = ix.pack_texts(ascii_texts)
    assert keep_ids is None and offsets[0] == 0 and offsets[-1] == len(blob) and len(offsets) == len(ascii_texts) + 1
    blob, offsets, keep_ids = ix.pack_texts(uni_texts)
    assert keep_ids.tolist() == [0, 2, 4] and blob.decode("utf-8").startswith("café kelvin k")
    # explicit ids travel with the kept records
    idx = ma.create_index()
    idx.add_texts(["a b", " ", "c"], ids=np.array([10, 11, 12]))
    assert ix.vector_to_array(idx.id_map).tolist() == [10, 12]
    # the stable-hash rebuild is this one call
    idx2 = ma.rebuild_index_from_texts(ascii_texts, hash_fn=None)
    assert ix.vector_to_array(idx2.id_map).tolist() == [0, 3, 5, 6]
