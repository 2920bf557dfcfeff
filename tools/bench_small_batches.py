#!/usr/bin/env python3
"""Device time per batch for small batches (2..256 queries) on the tensor-core path, with the queries-resident form
(gemm_rows_form = 1, default) and with the 256 x 256 form (0), next to one single-query scan.
usage: python tools/bench_small_batches.py [--rows 10000000] [--d 768] [--metric 0] [--k 10]"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import c99_vectordb_b200 as m
from oracle import oracle


def timed(idx, q, k, reps=6):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D, I = idx.search_device(q, k)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), I.clone()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--metric", type=int, default=0)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    idx = m.IndexFlat(a.d, a.metric)
    idx.add_synthetic(a.rows, 1234)
    q1 = torch.from_numpy(oracle.synth_rows(1, a.d, 99)).cuda()
    scan_ms, _ = timed(idx, q1, a.k)
    shadow_gb = a.rows * ((a.d + (2 if a.metric else 0) + 63) // 64 * 64) * 2 / 1e9
    for nq in (2, 4, 8, 16, 32, 64, 96, 128, 192, 256):
        q = torch.from_numpy(oracle.synth_rows(nq, a.d, 5678)).cuda()
        out = {"rows": a.rows, "d": a.d, "metric": "l2" if a.metric else "ip", "k": a.k, "nq": nq, "single_query_scan_ms": round(scan_ms, 3)}
        ids = {}
        for form in (1, 0):
            idx.set_option("gemm_rows_form", form)
            ms, I = timed(idx, q, a.k)
            ids[form] = I
            key = "rows_form" if form else "tile256_form"
            out[key + "_ms"] = round(ms, 3)
            out[key + "_used"] = bool(idx.get_option("stat_gemm_rows_form")) if form else False
            out[key + "_emit_ms"] = idx.get_option("stat_gemm_pass2_us") / 1e3
            out[key + "_uncertified"] = idx.get_option("stat_gemm_fallbacks")
        out["shadow_GBps_rows_form_emit"] = round(shadow_gb / (out["rows_form_emit_ms"] * 1e-3)) if out["rows_form_emit_ms"] else None
        out["ids_equal"] = bool((ids[0] == ids[1]).all().item())
        out["qps_rows_form"] = round(nq / (out["rows_form_ms"] * 1e-3))
        print(json.dumps(out), flush=True)
        assert out["ids_equal"]


if __name__ == "__main__":
    main()
