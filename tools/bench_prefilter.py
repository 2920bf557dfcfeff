#!/usr/bin/env python3
"""Single-query search with and without option `prefilter` (bf16 shadow ranked first, exact fp32 re-rank, certificate):
device time per search (CUDA events around back-to-back searches with distinct queries), ids equal, certificates held.
usage: python tools/bench_prefilter.py [--rows 10000000] [--d 768] [--metric 0] [--k 10] [--steps 50]"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import c99_vectordb_b200 as m
from oracle import oracle


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--metric", type=int, default=0)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--normalize", type=int, default=0)
    a = ap.parse_args()
    idx = m.IndexFlat(a.d, a.metric, normalize=bool(a.normalize))
    idx.add_synthetic(a.rows, 1234)
    qs = torch.from_numpy(oracle.synth_rows(a.steps + 5, a.d, 5678)).cuda()
    out = {}
    res = {}
    for pf in (0, 1, 0, 1):
        idx.set_option("prefilter", pf)
        for s in range(5):
            idx.search_device(qs[s:s + 1], a.k)
        torch.cuda.synchronize()
        fb0 = idx.get_option("stat_prefilter_fallbacks")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ids = []
        e0.record()
        for s in range(a.steps):
            D, I = idx.search_device(qs[5 + s:6 + s], a.k)
            ids.append(I.clone())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out.setdefault(pf, []).append(ms)
        res[pf] = torch.stack(ids).cpu().numpy()
        if pf:
            out["fallbacks"] = idx.get_option("stat_prefilter_fallbacks") - fb0
    elem_bytes = a.rows * a.d
    line = {"rows": a.rows, "d": a.d, "metric": "l2" if a.metric else "ip", "k": a.k, "steps": a.steps,
            "fp32_scan_ms": round(min(out[0]), 4), "prefilter_ms": round(min(out[1]), 4),
            "fp32_scan_qps": round(1e3 / min(out[0]), 1), "prefilter_qps": round(1e3 / min(out[1]), 1),
            "fp32_scan_GBps": round(elem_bytes * 4 / min(out[0]) / 1e6, 1), "shadow_scan_GBps": round(elem_bytes * 2 / min(out[1]) / 1e6, 1),
            "uncertified_of_steps": out["fallbacks"], "ids_identical": bool((res[0] == res[1]).all())}
    print(json.dumps(line), flush=True)
    assert line["ids_identical"]


if __name__ == "__main__":
    main()
