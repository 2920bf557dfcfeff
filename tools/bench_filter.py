"""Filtered recall by record id: building the row mask on the host (ids D2H + np.isin, the
first version) vs on the device (b200_index_search_ids_allowed).  10M x 384 L2, k = 10, one query."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import c99_vectordb_b200 as m
from oracle import oracle

n, d = 10_000_000, 384
for kind in ("dense", "sparse"):
    idx = m.IndexIDMap2(m.IndexFlat(d, 1))
    idx.index.add_synthetic(n, 1234, with_ids=True)   # ids = row positions (memo's record ids)
    if kind == "sparse":                              # rebuild with ids spread over 2^50
        rows = idx.index.reconstruct_n(0, 1)          # touch; keep the rows, replace the ids on the device is not exposed:
        idx = m.IndexIDMap2(m.IndexFlat(d, 1))
        ids = np.arange(n, dtype=np.int64) * 100_000_007
        step = 1_000_000
        for lo in range(0, n, step):
            idx.add_with_ids(oracle.synth_rows(step, d, 1234, first_row=lo), ids[lo:lo + step])
    all_ids = m.vector_to_array(idx.id_map)
    q = oracle.synth_rows(1, d, 5678)
    rng = np.random.default_rng(1)
    for frac in (0.001, 0.01, 0.5):
        allowed = all_ids[rng.random(n) < frac]
        rng.shuffle(allowed)
        t_dev, t_host = [], []
        for rep in range(4):
            t0 = time.perf_counter(); D1, I1 = idx.search(q, 10, ids_allowed=allowed); t_dev.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            mask = np.isin(m.vector_to_array(idx.id_map), allowed)
            D2, I2 = idx.search(q, 10, row_mask=mask)
            t_host.append(time.perf_counter() - t0)
            assert np.array_equal(I1, I2) and np.array_equal(D1, D2)
        print(json.dumps(dict(ids=kind, n=n, d=d, allowed=int(allowed.size), device_mask_ms=round(1e3 * sorted(t_dev[1:])[1], 2),
                              host_mask_ms=round(1e3 * sorted(t_host[1:])[1], 2))), flush=True)
    del idx
