#!/usr/bin/env python3
"""A/B of the two large-batch forms of K3: queries resident in tensor memory (gemm_ts.cuh, option gemm_ts = 1)
against the 256 x 256 smem-operand form (gemm_topk.cuh, gemm_ts = 0).

  python tools/bench_ts.py [--skip-small] [--big ROWS] [--nq NQ] [--reps R] > profiles/r2_gemm_ts_ab.jsonl

Small cases: both forms must return the same ids and distances as each other (and the candidate totals must agree:
the two forms compute the same bf16 products); every case prints one JSON line.  Big case (config 2 by default):
emit-pass / whole-batch times of both forms, interleaved."""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np
import torch

import c99_vectordb_b200 as m
from c99_vectordb_b200 import _cabi

STATS = ("stat_gemm_used", "stat_gemm_ts", "stat_gemm_rows_form", "stat_gemm_fallbacks", "stat_gemm_cand_total", "stat_gemm_pass1_us",
         "stat_gemm_pass2_us", "stat_gemm_rerank_us", "stat_gemm_streamed")


def synth(n, d, seed):
    out = torch.empty((n, d), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.load().b200_synth_rows_dev(out.data_ptr(), n, d, seed, 0, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out


def small_case(n, d, nq, k, metric, masked=False, streamed=False, chunk_tiles=0):
    idx = m.IndexFlat(d, metric)
    idx.set_option("gemm_min_nq", 2)
    idx.set_option("gemm_rows_form", 0)
    if streamed:
        idx.set_option("gemm_shadow_max_rows", 65536)
    if chunk_tiles:
        idx.set_option("gemm_chunk_tiles", chunk_tiles)
    idx.add_synthetic(n, 1234)
    q = synth(nq, d, 777)
    mask = np.random.default_rng(5).random(n) < 0.5 if masked else None
    out = {}
    for ts in (0, 1):
        idx.set_option("gemm_ts", ts)
        if masked:
            D, I = idx.search(q.cpu().numpy(), k, row_mask=mask)
        else:
            D, I = idx.search_device(q, k)
            torch.cuda.synchronize()
            D, I = D.cpu().numpy().copy(), I.cpu().numpy().copy()
        out[ts] = (D, I, {s: idx.get_option(s) for s in STATS})
    same = bool((out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all())
    line = {"case": f"{n}x{d} nq={nq} k={k} metric={'ip' if metric == 0 else 'l2'}" + (" masked" if masked else "") + (" streamed" if streamed else "")
            + (f" chunk_tiles={chunk_tiles}" if chunk_tiles else ""),
            "results_identical": same, "smem_form": out[0][2], "tmem_form": out[1][2],
            "ok": same and out[1][2]["stat_gemm_ts"] == 1 and out[0][2]["stat_gemm_ts"] == 0
            and out[0][2]["stat_gemm_cand_total"] == out[1][2]["stat_gemm_cand_total"]
            and out[1][2]["stat_gemm_fallbacks"] == out[0][2]["stat_gemm_fallbacks"]}
    print(json.dumps(line), flush=True)
    idx.close()
    return line["ok"]


def big_case(n, d, nq, k, reps, chunk_sweep):
    """Both forms interleaved; SM clock sampled (NVML) while each batch runs, so the times can be read per clock."""
    from bench import ClockSampler

    idx = m.IndexFlat(d, 0)
    idx.add_synthetic(n, 1234)
    q = synth(nq, d, 5678)
    variants = [(0, 0), (1, 0)] + [(1, c) for c in chunk_sweep]
    res = {v: [] for v in variants}
    ids = {}
    for r in range(reps + 1):
        for v in variants:
            ts, ct = v
            idx.set_option("gemm_ts", ts)
            idx.set_option("gemm_chunk_tiles", ct)
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
        ts, ct = v
        rs = sorted(res[v], key=lambda x: x["stat_gemm_pass2_us"])
        med = rs[len(rs) // 2]
        mhz = med["sm_mhz"]
        flops = 2.0 * nq * n * d
        line = {"case": f"{n}x{d} nq={nq} k={k} ip", "form": "tmem" if ts else "smem", "chunk_tiles": ct or "auto", "ts": med["stat_gemm_ts"],
                "emit_ms_median": med["stat_gemm_pass2_us"] / 1e3, "emit_ms_all": [x["stat_gemm_pass2_us"] / 1e3 for x in rs],
                "emit_tflops": flops / (med["stat_gemm_pass2_us"] * 1e-6) / 1e12,
                "batch_ms_median": sorted(x["batch_ms"] for x in rs)[len(rs) // 2], "pass1_ms": med["stat_gemm_pass1_us"] / 1e3,
                "rerank_ms": med["stat_gemm_rerank_us"] / 1e3, "fallbacks": med["stat_gemm_fallbacks"],
                "cand_total": med["stat_gemm_cand_total"], "sm_mhz_during_batch": mhz, "clock_reasons": med["reasons"]}
        if mhz:
            # dense bf16: 8192 flop per clock per SM (2.25 PFLOP/s nominal at 1.86 GHz over 148 SMs)
            line["tensor_pipe_frac_at_that_clock"] = flops / (med["stat_gemm_pass2_us"] * 1e-6) / (148 * 8192 * mhz * 1e6)
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
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--big", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--chunk-sweep", type=int, nargs="*", default=[])
    a = ap.parse_args()
    ok = True
    if not a.skip_small:
        ok &= small_case(200_000, 768, 300, 100, 0)                  # NT = 64, kps = 4
        ok &= small_case(131_072 + 77, 384, 700, 10, 1)              # L2: kpad 448 -> NT = 128, kps = 1; ragged last tile
        ok &= small_case(150_000, 256, 513, 10, 0, chunk_tiles=5)    # NT = 128, kps = 2; many units per CTA pair
        ok &= small_case(120_000, 640, 260, 20, 0, masked=True)      # NT = 64, kps = 2, filtered
        ok &= small_case(300_000, 768, 400, 10, 0, streamed=True)    # streamed shadow halves
        ok &= small_case(90_000, 100, 257, 10, 0)                    # kpad 128
        print(json.dumps({"small_cases_ok": bool(ok)}), flush=True)
    if a.big > 0:
        big_case(a.big, a.d, a.nq, a.k, a.reps, a.chunk_sweep)
    return 0 if ok else 1


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(json.dumps({"wall_s": round(time.time() - t0, 1)}))
    sys.exit(rc)
