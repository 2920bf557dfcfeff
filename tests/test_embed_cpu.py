"""The native stable-hash embedder (host code in the C-ABI library, no GPU needed) reproduces the
reference's embed_text_hash under PYTHONHASHSEED=0 bit for bit (tests/golden/embed.npz was produced
by the unmodified memo_cli.py, see tests/golden/make_golden.py)."""
import subprocess
import sys

import numpy as np

from conftest import GOLDEN
from c99_vectordb_b200 import memo_adapter as ma


def test_stable_hash_equals_cpython_seed0():
    tokens = ["a", "hello", "world_1", "peanuts", "0123456789abcdef", "x" * 33, "_", "z9"]
    code = "import sys; print([hash(t) for t in %r])" % (tokens,)
    out = subprocess.run([sys.executable, "-c", code], env={"PYTHONHASHSEED": "0", "PATH": "/usr/bin:/bin"},
                         capture_output=True, text=True, check=True).stdout
    assert [ma.stable_hash(t) for t in tokens] == eval(out)


def test_bulk_embedder_matches_reference_vectors():
    g = np.load(GOLDEN / "embed.npz", allow_pickle=True)
    texts = g["texts"].tolist()
    raw = ma.embed_texts_stable(texts)
    assert raw.shape == (len(texts), ma.DIM) and raw.dtype == np.float32
    # the same buckets as the Python token loop with the stable hash injected
    slow = ma.embed_texts(texts, hash_fn=ma.stable_hash)
    np.testing.assert_array_equal(raw, slow)
    # and, once normalised as memo_cli.normalize does (numpy), the reference's vectors exactly
    ref = g["vectors"]
    for r, want in zip(raw, ref):
        n = np.linalg.norm(r)
        got = np.zeros_like(r) if n <= 1e-8 else r / n
        np.testing.assert_array_equal(got.astype(np.float32), want)


def test_tokeniser_edge_cases():
    v = ma.embed_texts_stable(["", "!!! ...", "Hello hello HELLO", "snake_case x1 x1", "café K"])
    assert not v[0].any() and not v[1].any()
    assert np.abs(v[2]).sum() == 3 and np.count_nonzero(v[2]) == 1  # one token three times
    # 'café' tokenises as 'caf' (é is not [a-z0-9_]); the Kelvin sign lower-cases to ASCII 'k' in Python
    slow = ma.embed_texts(["café K"], hash_fn=ma.stable_hash)
    np.testing.assert_array_equal(v[4:5], slow)
