"""K6 — the hashing-trick embedder + K1 on the device (b200_index_add_texts) against the reference's own vectors
(tests/golden/embed.npz, recorded from the unmodified embed_text_hash under PYTHONHASHSEED=0) and against the host
embedder (csrc/embed.cu, itself pinned to CPython's hash in tests/test_embed_cpu.py): bit-exact rows, ids, blank
skipping, chunk boundaries."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ma(gpu):
    from c99_vectordb_b200 import memo_adapter

    return memo_adapter


def _host_rows(ma, texts, dim):
    raw = ma.embed_texts_stable([t.lower() if not t.isascii() else t for t in texts], dim)
    # the host embedder expects lower-cased text; embed_texts_stable lower-cases ASCII itself
    out = np.zeros_like(raw)
    for i, r in enumerate(raw):
        n = np.linalg.norm(r)  # memo_cli.normalize: numpy norm, true division, <= 1e-8 -> zeros
        out[i] = 0 if n <= 1e-8 else r / n
    return out.astype(np.float32)


def test_golden_vectors_of_the_reference(ma):
    g = np.load(GOLDEN / "embed.npz", allow_pickle=True)
    texts, ref = g["texts"].tolist(), g["vectors"]
    idx = ma.create_index()
    added = idx.add_texts(texts)
    keep = [i for i, t in enumerate(texts) if not ma.is_blank_body(t)]
    assert added == len(keep)
    np.testing.assert_array_equal(ma._ix.vector_to_array(idx.id_map), np.asarray(keep, dtype=np.int64))
    np.testing.assert_array_equal(idx.index.reconstruct_n(0, added), ref[keep])


def _corpus(n, seed):
    rng = np.random.default_rng(seed)
    words = ["a", "B2", "peanuts", "ALLERGY", "wifi_password", "x" * 8, "y" * 9, "z" * 16, "w" * 17, "q" * 40,
             "Tok3n", "_", "__init__", "0", "1234567", "12345678", "UPPER", "MiXeD"]
    seps = [" ", "  ", "\t", ", ", ". ", " - ", "\n", "!? ", "/", "'s "]
    out = []
    for i in range(n):
        r = i % 97
        if r == 0:
            out.append("")
        elif r == 1:
            out.append(" \t\n\x0b\x0c\r\x1c\x1f ")
        elif r == 2:
            out.append("?!... --- ###")  # no token: zero vector, but NOT blank
        else:
            k = int(rng.integers(1, 80))
            parts = []
            for _ in range(k):
                parts.append(words[int(rng.integers(len(words)))] + (str(int(rng.integers(1000))) if rng.random() < 0.3 else ""))
                parts.append(seps[int(rng.integers(len(seps)))])
            out.append("".join(parts))
    return out


@pytest.mark.parametrize("dim,store", [(384, "f32"), (1024, "bf16"), (100, "f32"), (37, "f32")])
def test_random_corpus_equals_host_embedder(ma, dim, store):
    import c99_vectordb_b200 as m

    texts = _corpus(30_000, 5)
    texts[777] = "long " + " ".join(f"tok{i}" for i in range(30_000))  # one ~200 KB record
    idx = m.IndexIDMap2(m.IndexFlat(dim, 1, store=store))
    added = idx.add_texts(texts)
    keep = [i for i, t in enumerate(texts) if not ma.is_blank_body(t)]
    assert added == len(keep) == idx.ntotal
    np.testing.assert_array_equal(m.vector_to_array(idx.id_map), np.asarray(keep, dtype=np.int64))
    want = _host_rows(ma, [texts[i] for i in keep], dim)
    got = idx.index.reconstruct_n(0, added)
    if store == "bf16":
        from oracle import oracle

        want = oracle.round_bf16(want)
    np.testing.assert_array_equal(got, want)
    # a second call appends after the first (row base continues), explicit ids
    more = ["appended one", "  ", "appended TWO two"]
    assert idx.add_texts(more, ids=np.array([10**6, 10**6 + 1, 10**6 + 2])) == 2
    assert m.vector_to_array(idx.id_map)[-2:].tolist() == [10**6, 10**6 + 2]


def test_non_ascii_records(ma):
    texts = ["café Kelvin K", "  ", "Straße ÄÖÜ groß", "", "İstanbul ǅ mixed", "plain ascii"]
    idx = ma.create_index()
    added = idx.add_texts(texts)
    keep = [i for i, t in enumerate(texts) if not ma.is_blank_body(t)]
    assert added == len(keep)
    slow = ma.embed_texts([texts[i] for i in keep], hash_fn=ma.stable_hash)
    want = np.zeros_like(slow)
    for i, r in enumerate(slow):
        n = np.linalg.norm(r)
        want[i] = 0 if n <= 1e-8 else r / n
    np.testing.assert_array_equal(idx.index.reconstruct_n(0, added), want.astype(np.float32))


def test_chunk_boundaries_and_rebuild_search(ma):
    """More records than one upload chunk holds (2^20): positions and ids stay in record order across chunks; the
    rebuilt index answers like the reference-shaped rebuild from host vectors."""
    n = (1 << 20) + 5000
    texts = [f"rec{i} w{i % 13} v{i % 7}" if i % 1000 else " " for i in range(n)]
    idx = ma.rebuild_index_from_texts(texts, hash_fn=None)
    keep = np.asarray([i for i in range(n) if i % 1000], dtype=np.int64)
    assert idx.ntotal == keep.size
    np.testing.assert_array_equal(ma._ix.vector_to_array(idx.id_map), keep)
    for probe in (0, 1, 999_000, keep.size - 1):  # rows around the chunk edge and at the ends
        want = _host_rows(ma, [texts[int(keep[probe])]], ma.DIM)[0]
        np.testing.assert_array_equal(idx.index.reconstruct(probe), want)
    target = 1048577  # a record of the second chunk
    q = ma.embed_text_hash(texts[target], hash_fn=ma.stable_hash)
    res = ma.search_all(idx, q, k=3)
    # 384 buckets: other records collide with this one's "rec…" token, so the exact match is a tie class whose
    # smallest id comes first; every member of it shares the record's w / v tokens
    assert res[0].score < 1e-6 and res[0].doc_id % 13 == target % 13 and res[0].doc_id % 7 == target % 7
    row = int(np.searchsorted(keep, target))
    np.testing.assert_array_equal(idx.index.reconstruct(row), q)
