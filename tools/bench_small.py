"""memo-sized databases (d = 384, memo_cli.py:17): single-query latency at k = 10 and k = ntotal (search_all,
memo_cli.py:288-298), device time (CUDA events) and through the host API (numpy in, numpy out)."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import c99_vectordb_b200 as m
from oracle import oracle

d = 384
for n in (1_000, 4_000, 10_000, 100_000, 1_000_000):
    idx = m.IndexIDMap2(m.IndexHNSWFlat(d, 32))          # what memo's create_index builds: flat L2 behind the shim
    idx.index.add_synthetic(n, 1234, with_ids=True)
    q = oracle.synth_rows(1, d, 5678)
    qt = torch.from_numpy(q).cuda()
    out = dict(n=n, d=d)
    for name, k in (("k10", 10), ("kall", n)):
        D = torch.empty((1, k), dtype=torch.float32, device="cuda"); I = torch.empty((1, k), dtype=torch.int64, device="cuda")
        for _ in range(5): idx.search_device(qt, k, D=D, I=I)
        torch.cuda.synchronize()
        ts = []
        for _ in range(50):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); idx.search_device(qt, k, D=D, I=I); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        hs = []
        for _ in range(55):
            t0 = time.perf_counter(); idx.search(q, k); hs.append((time.perf_counter() - t0) * 1e6)
        out[f"{name}_device_us"] = round(sorted(ts)[25], 1)
        out[f"{name}_host_api_us"] = round(sorted(hs[5:])[25], 1)
    print(json.dumps(out), flush=True)
