#!/usr/bin/env python3
"""Device time of one large batch on the tensor-core path (K3, config 2 by default) with the SM clock sampled while
it runs, and the library GEMM of the same shape on the same box for scale.

  python tools/bench_batched.py [--rows N] [--d D] [--nq NQ] [--k K] [--reps R] [--chunk-sweep C ...]

One JSON line per variant (chunk_tiles = auto and every value of --chunk-sweep), then the cuBLAS line.
(The first version of this tool, tools/bench_ts.py in commit bec616f, also A/B-tested a kernel that kept the queries in
tensor memory; its results are profiles/r2_gemm_ts_experiment.jsonl.)"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

import c99_vectordb_b200 as m
from c99_vectordb_b200 import _cabi

STATS = ("stat_gemm_used", "stat_gemm_rows_form", "stat_gemm_fallbacks", "stat_gemm_cand_total", "stat_gemm_pass1_us",
         "stat_gemm_pass2_us", "stat_gemm_rerank_us", "stat_gemm_streamed")


def synth(n, d, seed):
    out = torch.empty((n, d), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.load().b200_synth_rows_dev(out.data_ptr(), n, d, seed, 0, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out


def big_case(n, d, nq, k, reps, chunk_sweep):
    """SM clock sampled (NVML, a lagging reading) while each batch runs."""
    from bench import ClockSampler

    idx = m.IndexFlat(d, 0)
    idx.add_synthetic(n, 1234)
    q = synth(nq, d, 5678)
    variants = [0] + list(chunk_sweep)
    res = {v: [] for v in variants}
    ids = {}
    for r in range(reps + 1):
        for v in variants:
            idx.set_option("gemm_chunk_tiles", v)
            sam = ClockSampler(0)
            if r:
                sam.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            D, I = idx.search_device(q, k)
            e1.record()
            torch.cuda.synchronize()
            if r == 0:
                ids[v] = I.cpu().numpy().copy()
                continue  # warm-up
            clk = sam.stop()
            res[v].append({"batch_ms": e0.elapsed_time(e1), "sm_mhz": clk.get("sm_mhz"), "reasons": clk.get("reasons"),
                           **{s: idx.get_option(s) for s in STATS}})
    for v in variants:
        ct = v
        rs = sorted(res[v], key=lambda x: x["stat_gemm_pass2_us"])
        med = rs[len(rs) // 2]
        mhz = med["sm_mhz"]
        flops = 2.0 * nq * n * d
        line = {"case": f"{n}x{d} nq={nq} k={k} ip", "chunk_tiles": ct or "auto",
                "emit_ms_median": med["stat_gemm_pass2_us"] / 1e3, "emit_ms_all": [x["stat_gemm_pass2_us"] / 1e3 for x in rs],
                "emit_tflops": flops / (med["stat_gemm_pass2_us"] * 1e-6) / 1e12,
                "batch_ms_median": sorted(x["batch_ms"] for x in rs)[len(rs) // 2], "pass1_ms": med["stat_gemm_pass1_us"] / 1e3,
                "rerank_ms": med["stat_gemm_rerank_us"] / 1e3, "fallbacks": med["stat_gemm_fallbacks"],
                "cand_total": med["stat_gemm_cand_total"], "sm_mhz_during_batch": mhz, "clock_reasons": med["reasons"]}
        print(json.dumps(line), flush=True)
    print(json.dumps({"case": "big", "ids_identical": bool(all((ids[v] == ids[variants[0]]).all() for v in variants))}), flush=True)
    idx.close()
    # the library GEMM of the same shape on the same box, for scale: Q[nq, d] x rows[chunk, d]^T in bf16 (cuBLAS through
    # torch), bf16 output that is never read — a quarter-million rows per call so that the output stays in reach
    rows = 262144
    a = torch.randn(nq, d, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(rows, d, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(nq, rows, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, b.t(), out=out)
    torch.cuda.synchronize()
    calls = 20
    sam = ClockSampler(0)
    sam.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        torch.matmul(a, b.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    clk = sam.stop()
    ms = e0.elapsed_time(e1) / calls
    print(json.dumps({"case": f"cuBLAS bf16 {nq}x{d} @ {d}x{rows} -> bf16 (torch.matmul), {calls} calls", "ms_per_call": ms,
                      "tflops": 2.0 * nq * rows * d / (ms * 1e-3) / 1e12, "equivalent_ms_for_the_whole_database": ms * n / rows,
                      "sm_mhz": clk.get("sm_mhz"), "note": "writes 2 bytes per score (5.2 GB per call); K3 writes none"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunk-sweep", type=int, nargs="*", default=[])
    a = ap.parse_args()
    big_case(a.rows, a.d, a.nq, a.k, a.reps, a.chunk_sweep)
    return 0


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(json.dumps({"wall_s": round(time.time() - t0, 1)}))
    sys.exit(rc)
