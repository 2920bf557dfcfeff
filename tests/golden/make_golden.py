"""Generate tests/golden/*.npz from the REFERENCE's own Python functions.

Run in the build container only (needs /root/reference):
    PYTHONHASHSEED=0 python tests/golden/make_golden.py

memo_cli.py cannot be imported without a `faiss` module (memo_cli.py:13) and faiss-cpu is absent
here, so tests/stub_faiss_oracle.py (exact flat semantics from oracle/) is injected as `faiss`.
What gets pinned:
  normalize.npz      memo_cli.normalize()            — pure numpy, independent of the stub
  embed.npz          memo_cli.embed_text_hash()      — pure numpy + hash() under PYTHONHASHSEED=0
  adapter.npz        memo_cli.rebuild_index_from_texts / get_existing_ids / search_all results —
                     the reference's adapter logic (blank skipping, sparse ids, k = ntotal,
                     id < 0 dropping) over the oracle's flat L2 ranking
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0"

import stub_faiss_oracle  # noqa: E402

sys.modules["faiss"] = stub_faiss_oracle
sys.path.insert(0, "/root/reference")
import memo_cli  # noqa: E402  (the reference, unmodified)

OUT = Path(__file__).resolve().parent
rng = np.random.default_rng(20261018)

# ---- normalize ---------------------------------------------------------------------------------
cases = [rng.standard_normal(384).astype(np.float32) for _ in range(8)]
cases += [rng.uniform(-1, 1, 768).astype(np.float32), rng.uniform(-1, 1, 1024).astype(np.float32)]
cases += [np.zeros(384, np.float32), np.full(384, 1e-12, np.float32), np.full(16, 3e-10, np.float32),
          np.array([3.0, 4.0], np.float32), np.array([1e-8, 0, 0, 0], np.float32),
          np.array([1.0000001e-8, 0, 0, 0], np.float32), rng.uniform(-1, 1, 7).astype(np.float32)]
norm = {}
for i, v in enumerate(cases):
    norm[f"in_{i}"] = v
    norm[f"out_{i}"] = memo_cli.normalize(v).astype(np.float32)
np.savez(OUT / "normalize.npz", n=len(cases), **norm)

# ---- embed -------------------------------------------------------------------------------------
texts = [
    "My daughter's name is Sarah and she is allergic to peanuts.",
    "The wifi password at the office is hunter2",
    "Remember to rotate the API keys every 90 days",
    "peanuts allergies",
    "   Whitespace\t\tand\nnewlines   collapse ",
    "UPPER lower MiXeD snake_case 12345",
    "",
    "!!! ??? ...",
    "flat index search on B200 with fused top-k",
    "the quick brown fox jumps over the lazy dog " * 5,
]
words = ("alpha beta gamma delta epsilon zeta eta theta iota kappa lambda mu nu xi omicron pi rho sigma tau "
         "upsilon phi chi psi omega memo vector index recall save yaml gpu kernel warp tile shard merge").split()
for i in range(54):
    n = int(rng.integers(3, 24))
    texts.append(" ".join(rng.choice(words, n)))
emb = np.stack([memo_cli.embed_text_hash(t) for t in texts]).astype(np.float32)
np.savez(OUT / "embed.npz", texts=np.array(texts, dtype=object), vectors=emb, allow_pickle=True)

# ---- adapter ------------------------------------------------------------------------------------
records = list(texts)
records[3] = None            # blank records are skipped -> sparse ids (memo_cli.py:278-280)
records[6] = ""
records[7] = "   "
index = memo_cli.rebuild_index_from_texts(records, verbose=False)
existing = sorted(memo_cli.get_existing_ids(index))
queries = ["peanuts allergies", "wifi password", "gpu kernel tile", "omega alpha", "nothing matches zzz"]
qvecs = np.stack([memo_cli.embed_text_hash(q) for q in queries]).astype(np.float32)
res_ids, res_scores = [], []
for qv in qvecs:
    rs = memo_cli.search_all(index, qv)
    res_ids.append(np.array([r.doc_id for r in rs], dtype=np.int64))
    res_scores.append(np.array([r.score for r in rs], dtype=np.float32))
kept = [i for i, t in enumerate(records) if not memo_cli.is_blank_body(t or "")]
np.savez(OUT / "adapter.npz", records=np.array([r if r is not None else "\x00NONE" for r in records], dtype=object),
         kept=np.array(kept, dtype=np.int64), kept_vectors=emb[kept], existing_ids=np.array(existing, dtype=np.int64),
         queries=np.array(queries, dtype=object), qvecs=qvecs,
         res_ids=np.stack(res_ids), res_scores=np.stack(res_scores), allow_pickle=True)
print("wrote", [p.name for p in OUT.glob("*.npz")])

# ---- filtered recall: the reference's own command_recall --yaml --filter ----------
import contextlib  # noqa: E402
import io  # noqa: E402
import tempfile  # noqa: E402

import yaml  # noqa: E402

fr_texts = [t for t in texts[:40]]
fr_texts[5] = ""  # blank body: never shown (memo_cli.py:509)
fr_metas = []
for i in range(len(fr_texts)):
    if i % 7 == 3:
        fr_metas.append(None)  # records without metadata never match a filter (memo_cli.py:502-504)
    else:
        fr_metas.append({"priority": i % 5, "topic": ["gpu", "memo", "yaml"][i % 3], "tags": ["a", "b"] if i % 2 else ["c"]})
filters = [None, "priority: {$gte: 3}", "topic: gpu", "{$or: [{topic: memo}, {priority: 0}]}", "tags: {$contains: b}", "priority: 99"]
fr_queries = ["alpha beta gamma", "gpu kernel warp tile", "peanuts allergies", "omega psi chi"]
with tempfile.TemporaryDirectory() as td:
    idx_path, yaml_path = memo_cli.build_db_paths("frdb", td)
    memo_cli.save_yaml_tables(yaml_path, fr_texts, fr_metas)
    fr_index = memo_cli.rebuild_index_from_texts(fr_texts, verbose=False)
    stub_faiss_oracle.write_index(fr_index, str(idx_path))
    cases = []
    for qtext in fr_queries:
        for f in filters:
            for kk in (1, 3, 10):
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    rc = memo_cli.command_recall("frdb", qtext, kk, f, True, td)
                assert rc == 0
                res = yaml.safe_load(buf.getvalue())["results"]
                cases.append((qtext, f, kk, [r["id"] for r in res], [r["score"] for r in res]))
fr_kept = [i for i, t in enumerate(fr_texts) if not memo_cli.is_blank_body(t or "")]
eligible = {}
for f in filters:
    if f is None:
        eligible["None"] = fr_kept
    else:
        fd = memo_cli.parse_yaml_flow_map(f)
        eligible[f] = [i for i in fr_kept if fr_metas[i] and memo_cli.matches_filter(fr_metas[i], fd)]
np.savez(OUT / "recall_filter.npz",
         kept=np.array(fr_kept, dtype=np.int64),
         kept_vectors=np.stack([memo_cli.embed_text_hash(fr_texts[i]) for i in fr_kept]).astype(np.float32),
         queries=np.array(fr_queries, dtype=object),
         qvecs=np.stack([memo_cli.embed_text_hash(q) for q in fr_queries]).astype(np.float32),
         filters=np.array([str(f) for f in filters], dtype=object),
         eligible=np.array([np.array(eligible[str(f)], dtype=np.int64) for f in filters], dtype=object),
         cases=np.array([(q, str(f), kk, np.array(ids, dtype=np.int64), np.array(sc, dtype=np.float64)) for q, f, kk, ids, sc in cases], dtype=object),
         allow_pickle=True)
print("wrote recall_filter.npz with", len(cases), "cases")
