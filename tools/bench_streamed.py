#!/usr/bin/env python3
"""Batched queries on an index whose bf16 shadow does NOT fit next to the rows: 100M x 384 fp32 (153.6 GB) on one GPU.
The tensor-core path streams the rows through an L2-sized bf16 scratch (DESIGN.md 7.6 "streamed shadow"); before
round 2 such batches fell back to the register-blocked scan (8 queries per pass over the database).
Prints one JSON line per batch size with the device time per batch, the path taken and an ids check of 4 queries
against the exact scan kernel.  usage: python tools/bench_streamed.py [--rows 100000000] [--d 384] [--metric 1]"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import c99_vectordb_b200 as m
from oracle import oracle


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--d", type=int, default=384)
    ap.add_argument("--metric", type=int, default=1)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cap", type=int, default=0, help="force streaming with this many shadow rows (0 = only when it does not fit)")
    a = ap.parse_args()
    idx = m.IndexFlat(a.d, a.metric)
    idx.reserve(a.rows)
    idx.add_synthetic(a.rows, 1234)
    if a.cap:
        idx.set_option("gemm_shadow_max_rows", a.cap)
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    for nq in (8, 64, 256, 1024):
        q = torch.from_numpy(oracle.synth_rows(nq, a.d, 5678)).cuda()
        D = torch.empty((nq, a.k), dtype=torch.float32, device="cuda")
        I = torch.empty((nq, a.k), dtype=torch.int64, device="cuda")
        times = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx.search_device(q, a.k, D=D, I=I)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        st = {s: idx.get_option("stat_gemm_" + s) for s in ("used", "streamed", "fallbacks", "pass1_us", "pass2_us", "rerank_us")}
        # the exact scan as the referee for 4 of the queries
        idx.set_option("gemm_min_nq", 0)
        t0 = time.perf_counter()
        Ds, Is = idx.search_device(q[:4].contiguous(), a.k)
        torch.cuda.synchronize()
        scan4_ms = (time.perf_counter() - t0) * 1e3
        idx.set_option("gemm_min_nq", 2)
        same = bool((Is == I[:4]).all().item() and (Ds == D[:4]).all().item())
        print(json.dumps({"rows": a.rows, "d": a.d, "metric": "l2" if a.metric else "ip", "nq": nq, "k": a.k,
                          "ms_per_batch": round(min(times[1:]), 3), "first_ms": round(times[0], 3), "qps": round(nq / (min(times[1:]) * 1e-3)),
                          "path": st, "ids_and_distances_equal_exact_scan_4q": same, "exact_scan_4_queries_ms": round(scan4_ms, 2),
                          "free_GB_after_add": round(free / 1e9, 2)}), flush=True)
        assert same


if __name__ == "__main__":
    main()
