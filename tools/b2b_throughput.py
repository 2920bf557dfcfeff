#!/usr/bin/env python3
"""Back-to-back single-query throughput (what bench.py's `value` times) for a list of option sets — A/B of scan
launch parameters on one GPU.  One CUDA-event pair around `iters` searches on one stream, queries resident.

    python tools/b2b_throughput.py --rows 1250000 --d 768 --sets "" "scan_ctas_per_sm=2" "scan_pdl=0"
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_250_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--store", default="f32")
    ap.add_argument("--metric", default="ip")
    ap.add_argument("--sets", nargs="*", default=[""])
    a = ap.parse_args()
    import torch

    import c99_vectordb_b200 as m
    from c99_vectordb_b200 import _cabi

    L = _cabi.load()
    dev = torch.device("cuda", 0)
    q = torch.empty((a.iters, 1, a.d), dtype=torch.float32, device=dev)
    for s in range(a.iters):
        _cabi.check(L.b200_synth_rows_dev(q[s].data_ptr(), 1, a.d, 5678 + s, 0, 0, C.c_void_p(1)))
    idx = m.IndexFlat(a.d, 0 if a.metric == "ip" else 1, store=a.store)
    idx.add_synthetic(a.rows, 1234)
    idx.set_option("queries_stable", 1)
    defaults = {}
    D = torch.empty((1, a.k), dtype=torch.float32, device=dev)
    I = torch.empty((1, a.k), dtype=torch.int64, device=dev)
    bytes_ = a.rows * a.d * (4 if a.store == "f32" else 2)
    import random

    best, allms = {}, {}
    for rnd in range(a.rounds):  # interleaved rounds in shuffled order: drift and neighbours hit every set alike
        order = list(a.sets)
        random.Random(rnd).shuffle(order)
        for sset in order:
            opts = dict(kv.split("=") for kv in sset.split(",") if kv)
            for name, val in opts.items():
                defaults.setdefault(name, idx.get_option(name))
                idx.set_option(name, int(val))
            for s in range(5):
                idx.search_device(q[s], a.k, D=D, I=I)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(a.iters):
                idx.search_device(q[s], a.k, D=D, I=I)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            best[sset] = min(best.get(sset, 1e9), ms)
            allms.setdefault(sset, []).append(round(ms, 5))
            for name in opts:
                idx.set_option(name, defaults[name])
    for sset in a.sets:
        print(json.dumps({"rows": a.rows, "d": a.d, "store": a.store, "options": sset or "(defaults)", "ms_per_search": best[sset], "rounds_ms": allms[sset],
                          "gbs": bytes_ / best[sset] / 1e6, "qps": 1e3 / best[sset]}))


if __name__ == "__main__":
    main()
