"""Smallest end-to-end exercise of every kernel, for compute-sanitizer (one tool per gpurun call)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import c99_vectordb_b200 as m
from oracle import oracle

for metric in (0, 1):
    for store in ("f32", "bf16"):
        for variant in (1, 2):
            d, n = 72, 3001
            db, q = oracle.synth_rows(n, d, 1), oracle.synth_rows(5, d, 2)
            idx = m.IndexIDMap2(m.IndexFlat(d, metric, store=store, normalize=True))
            idx.index.set_option("scan_variant", variant)
            idx.add_with_ids(db, np.arange(n, dtype=np.int64) + 10)
            for k in (1, 10, 300, n + 5):
                idx.search(q, k)
            idx.search(q[:1], 7, row_mask=(np.arange(n) % 3 == 0))
big = m.IndexFlat(64, 0)
big.set_option("gemm_min_nq", 32)
big.add_synthetic(70000, 5)
for cg in (1, 2):
    big.set_option("gemm_cta_group", cg)
    D, I = big.search(oracle.synth_rows(40, 64, 6), 10)
    assert big.get_option("stat_gemm_used") == 1
x = oracle.synth_rows(100, 10, 3)
m.normalize_L2(x)
print("sanitize_small ok")
