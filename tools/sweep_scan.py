"""K2 variant / parameter sweep on one GPU (run under gpurun).  Times b200_index_search_dev with
CUDA events on torch's current stream and prints one JSON line per configuration.

    python tools/sweep_scan.py --n 10000000 --d 768 --metric ip --out gpurun_out/sweep_10Mx768.jsonl
"""
import argparse
import itertools
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import c99_vectordb_b200 as m  # noqa: E402


def time_config(idx, q, k, iters, warm=3):
    D = torch.empty((q.shape[0], k), dtype=torch.float32, device="cuda")
    I = torch.empty((q.shape[0], k), dtype=torch.int64, device="cuda")
    for _ in range(warm):
        idx.search_device(q, k, D=D, I=I)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        idx.search_device(q, k, D=D, I=I)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0], I.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--metric", default="ip")
    ap.add_argument("--store", default="f32")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=1)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--warps", action="store_true")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    a = ap.parse_args()
    metric = 0 if a.metric == "ip" else 1
    idx = m.IndexFlat(a.d, metric, store=a.store)
    t0 = time.time()
    idx.add_synthetic(a.n, seed=1234)
    gen_s = time.time() - t0
    q = torch.empty((a.nq, a.d), dtype=torch.float32, device="cuda")
    from c99_vectordb_b200 import _cabi
    import ctypes as C
    _cabi.check(_cabi.load().b200_synth_rows_dev(q.data_ptr(), a.nq, a.d, 5678, 0, 0, C.c_void_p(1)))
    torch.cuda.synchronize()
    bytes_per_query = a.n * a.d * (4 if a.store == "f32" else 2)
    peak = 6552.6
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    out = open(a.out, "a")
    configs = []
    if a.quick:
        configs += [dict(scan_variant=0), dict(scan_variant=1), dict(scan_variant=2)]
        if a.warps:
            configs += [dict(scan_variant=1, scan_warps=w) for w in (10, 12, 16)]
    else:
        for dyn, ef, tr in itertools.product((1, 0), (0, 1), (0, 8)):
            configs.append(dict(scan_variant=1, scan_tile_rows=tr, scan_l2_evict_first=ef, scan_dynamic_tiles=dyn))
        for dyn, cps in itertools.product((1, 0), (0, 3, 6)):
            configs.append(dict(scan_variant=2, scan_ctas_per_sm=cps, scan_dynamic_tiles=dyn))
        configs.append(dict(scan_variant=1, scan_dynamic_tiles=1, scan_claim_chunk=1))
        configs.append(dict(scan_variant=2, scan_dynamic_tiles=1, scan_claim_chunk=4))
        configs.append(dict(scan_variant=2, scan_dynamic_tiles=1, scan_claim_chunk=64))
    defaults = dict(scan_variant=0, scan_warps=8, scan_tile_rows=0, scan_stages=0, scan_l2_evict_first=0, scan_ctas_per_sm=0,
                    scan_dynamic_tiles=-1, scan_claim_chunk=0)
    ref_I = None
    results = {i: [] for i in range(len(configs))}
    for rnd in range(a.rounds):  # interleaved rounds: drift over the call hits every configuration alike
        for ci, cfg in enumerate(configs):
            for kname, v in {**defaults, **cfg}.items():
                idx.set_option(kname, v)
            try:
                p50, best, I = time_config(idx, q, a.k, a.iters)
            except Exception as e:
                results[ci].append(("error", str(e)))
                continue
            if ref_I is None:
                ref_I = I
            results[ci].append((p50, best, bool((I == ref_I).all())))
    for ci, cfg in enumerate(configs):
        good = [r for r in results[ci] if r[0] != "error"]
        if not good:
            rec = dict(cfg=cfg, error=results[ci][0][1])
        else:
            p50s = sorted(r[0] for r in good)
            med = p50s[len(p50s) // 2]
            gbs = bytes_per_query / (med * 1e-3) / 1e9
            rec = dict(n=a.n, d=a.d, metric=a.metric, store=a.store, k=a.k, nq=a.nq, cfg=cfg, p50_ms=round(med, 4),
                       min_ms=round(min(r[1] for r in good), 4), rounds=[round(x, 4) for x in p50s], gbs=round(gbs, 1),
                       frac=round(gbs / peak, 4), ids_consistent=all(r[2] for r in good))
        print(json.dumps(rec)); out.write(json.dumps(rec) + "\n"); out.flush()


if __name__ == "__main__":
    main()
