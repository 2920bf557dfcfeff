"""memo's search_all asks for k = ntotal (memo_cli.py:291): time the full-ranking path (score keys + stable
radix sort) on the device (search_device, no 12 B/row D2H) and through the host API."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import c99_vectordb_b200 as m
from oracle import oracle

for n, d in ((10_000, 384), (1_000_000, 384), (10_000_000, 384)):
    idx = m.IndexFlat(d, 1)
    idx.add_synthetic(n, 1234)
    q = oracle.synth_rows(1, d, 5678)
    qt = torch.from_numpy(q).cuda()
    D = torch.empty((1, n), dtype=torch.float32, device="cuda"); I = torch.empty((1, n), dtype=torch.int64, device="cuda")
    for _ in range(3): idx.search_device(qt, n, D=D, I=I)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); idx.search_device(qt, n, D=D, I=I); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    host = {}
    for rnd in range(3):  # interleaved A/B: results through the pinned ring vs plain copies into pageable memory
        for staged in (1, 0):
            idx.set_option("host_staged_results", staged)
            for it in range(3):
                t0 = time.perf_counter(); Dh, Ih = idx.search(q, n); ms = (time.perf_counter() - t0) * 1e3
                if it: host.setdefault(staged, []).append(ms)
    host_ms = sorted(host[1])[len(host[1]) // 2]
    host_plain_ms = sorted(host[0])[len(host[0]) // 2]
    ok = bool((np.diff(Dh[0]) >= 0).all()) and len(set(Ih[0].tolist())) == n
    print(json.dumps(dict(n=n, d=d, k=n, device_ms=round(sorted(ts)[5], 3), host_api_ms=round(host_ms, 2), host_api_plain_copy_ms=round(host_plain_ms, 2), sorted_and_complete=ok)), flush=True)
