"""Index rebuild (add) throughput — config 4's "full rebuild (add) from YAML-derived embeddings".
Replaces rebuild_index_from_texts' N single-row add_with_ids calls (memo_cli.py:276-282) with one bulk
add.  Measures (a) host-fed add: pageable numpy [n,d] fp32 -> resident rows (H2D + K1 normalise + store),
(b) device-fed add: rows already in HBM (K1 only), for fp32 and bf16 storage.  One JSON line each."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import c99_vectordb_b200 as m
from c99_vectordb_b200 import _cabi
import ctypes as C

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--d", type=int, default=1024)
ap.add_argument("--dev-n", type=int, default=10_000_000)
ap.add_argument("--host-only", action="store_true")
a = ap.parse_args()
rng = np.random.default_rng(0)
x = rng.standard_normal((a.n, a.d), dtype=np.float32)
ids = np.arange(a.n, dtype=np.int64)
for store in ("f32", "bf16"):
    for normalize in (False, True):
        best = None
        for rep in range(3):
            idx = m.IndexIDMap2(m.IndexFlat(a.d, 0, store=store, normalize=normalize))
            idx.reserve(a.n)
            t0 = time.perf_counter()
            idx.add_with_ids(x, ids)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            idx.index.close()
        print(json.dumps({"what": "host-fed add_with_ids", "n": a.n, "d": a.d, "store": store, "normalize": normalize,
                          "seconds": round(best, 4), "rows_per_s": round(a.n / best), "host_GBps": round(a.n * a.d * 4 / best / 1e9, 2)}), flush=True)
if a.host_only:
    sys.exit(0)
# device-fed: the rows are generated in HBM, then ingested by K1
src = torch.empty((a.dev_n, a.d), dtype=torch.float32, device="cuda")
_cabi.check(_cabi.load().b200_synth_rows_dev(src.data_ptr(), a.dev_n, a.d, 1234, 0, 0, C.c_void_p(1)))
torch.cuda.synchronize()
for store in ("f32", "bf16"):
    elem = 4 if store == "f32" else 2
    best = None
    for rep in range(3):
        idx = m.IndexFlat(a.d, 0, store=store, normalize=True)
        idx.reserve(a.dev_n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _cabi.check(_cabi.load().b200_index_add_dev(idx._h, src.data_ptr(), a.dev_n, None, 1))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        idx.close()
    byts = a.dev_n * a.d * (4 + elem)
    print(json.dumps({"what": "device-fed add (K1 normalise + store)", "n": a.dev_n, "d": a.d, "store": store, "seconds": round(best, 5),
                      "rows_per_s": round(a.dev_n / best), "hbm_GBps": round(byts / best / 1e9, 1), "algorithmic_bytes": byts}), flush=True)
