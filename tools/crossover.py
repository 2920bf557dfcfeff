"""K2 (scan, 8 queries per pass) vs K3 (tensor cores) for mid-size batches on one database."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import c99_vectordb_b200 as m
from c99_vectordb_b200 import _cabi
import ctypes as C

n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
metric = int(sys.argv[4]) if len(sys.argv) > 4 else 0
idx = m.IndexFlat(d, metric)
idx.add_synthetic(n, 1234)
for nq in (8, 16, 24, 32, 48, 64, 128, 256, 512):
    q = torch.empty((nq, d), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.load().b200_synth_rows_dev(q.data_ptr(), nq, d, 5678, 0, 0, C.c_void_p(1)))
    res = {}
    for name, min_nq in (("k2", 0), ("k3", 1)):
        idx.set_option("gemm_min_nq", min_nq)
        ts = []
        for it in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            D, I = idx.search_device(q, k)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        res[name] = (sorted(ts[1:])[len(ts[1:]) // 2] * 1e3, I.cpu().numpy())
        res[name + "_fb"] = idx.get_option("stat_gemm_fallbacks") if min_nq else 0
    same = bool((res["k2"][1] == res["k3"][1]).all())
    print(json.dumps(dict(n=n, d=d, k=k, metric=metric, nq=nq, k2_ms=round(res["k2"][0], 3), k3_ms=round(res["k3"][0], 3),
                          k3_fallbacks=res["k3_fb"], ids_identical=same)), flush=True)
