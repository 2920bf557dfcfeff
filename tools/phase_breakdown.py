#!/usr/bin/env python3
"""Per-phase timing of ONE scan launch (single query, fused top-k) from the kernel's own globaltimer stamps
(option scan_phase_stamps, b200_index_read_phase_stamps) next to the CUDA-event time of the same launch.

    python tools/phase_breakdown.py [--rows N] [--d D] [--k K] [--reps R] [--store f32|bf16] [--metric ip|l2]

Prints one JSON line per configuration: medians over R launches (ns unless named otherwise).
  t_first_cta_spread   latest CTA entry - earliest CTA entry
  prologue             CTA entry -> queries staged (mbarrier init, first tile requests, query staging), median over CTAs
  first_tile           CTA entry -> warp 0's first tile landed, median over CTAs
  scan_end_spread      latest - earliest "warp 0 finished scanning" over CTAs (the dynamic scheduler's tail)
  cta_reduce           warp 0 done -> this CTA's survivors written (includes waiting for the CTA's slowest warp)
  ticket               last CTA: survivors written -> final merge starts (fence + atomic ticket)
  final_merge          last CTA: merge of all CTAs' survivors + ids + D/I
  epilogue             last CTA: final merge done -> kernel end (exchange when sharded)
  kernel_span          earliest CTA entry -> last CTA end
  event_ms             CUDA events around the launch on its stream
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, nargs="+", default=[10_000_000, 1_250_000, 1_000_000])
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--store", default="f32")
    ap.add_argument("--metric", default="ip")
    a = ap.parse_args()
    import numpy as np
    import torch

    import c99_vectordb_b200 as m
    from c99_vectordb_b200 import _cabi

    L = _cabi.load()
    dev = torch.device("cuda", 0)
    for n in a.rows:
        idx = m.IndexFlat(a.d, 0 if a.metric == "ip" else 1, store=a.store)
        idx.add_synthetic(n, 1234)
        idx.set_option("scan_phase_stamps", 1)
        idx.set_option("scan_pdl", 0)  # anatomy of one launch alone (with PDL the kernel would wait for the stamp memset inside its prologue)
        q = torch.empty((a.reps + 3, 1, a.d), dtype=torch.float32, device=dev)
        for s in range(a.reps + 3):
            _cabi.check(L.b200_synth_rows_dev(q[s].data_ptr(), 1, a.d, 5678 + s, 0, 0, C.c_void_p(1)))
        torch.cuda.synchronize()
        buf = np.zeros(1024 * 8, dtype=np.uint64)
        rows = {k: [] for k in ("t_first_cta_spread", "prologue", "first_tile", "scan_end_spread", "scan_end_median",
                                "cta_reduce", "ticket", "final_merge", "epilogue", "kernel_span", "event_ms")}
        for s in range(a.reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            idx.search_device(q[s], a.k)
            e1.record()
            torch.cuda.synchronize()
            n_ctas = C.c_int64(0)
            _cabi.check(L.b200_index_read_phase_stamps(idx._h, buf.ctypes.data, buf.size, C.byref(n_ctas)))
            if s < 3:
                continue
            st = buf[: n_ctas.value * 8].reshape(-1, 8).astype(np.int64)
            t0 = st[:, 0].min()
            last = int(np.argmax(st[:, 6]))
            rows["t_first_cta_spread"].append(int(st[:, 0].max() - t0))
            rows["prologue"].append(int(np.median(st[:, 1] - st[:, 0])))
            ft = st[:, 7] - st[:, 0]
            rows["first_tile"].append(int(np.median(ft[st[:, 7] > 0])) if (st[:, 7] > 0).any() else 0)
            rows["scan_end_spread"].append(int(st[:, 2].max() - st[:, 2].min()))
            rows["scan_end_median"].append(int(np.median(st[:, 2]) - t0))
            rows["cta_reduce"].append(int(np.median(st[:, 3] - st[:, 2])))
            rows["ticket"].append(int(st[last, 4] - st[last, 3]))
            rows["final_merge"].append(int(st[last, 5] - st[last, 4]))
            rows["epilogue"].append(int(st[last, 6] - st[last, 5]))
            rows["kernel_span"].append(int(st[last, 6] - t0))
            rows["event_ms"].append(e0.elapsed_time(e1))
        out = {"rows": n, "d": a.d, "k": a.k, "store": a.store, "metric": a.metric, "reps": a.reps, "ctas": int(n_ctas.value),
               "scan_end_max_minus_t0": None}
        for k2, v in rows.items():
            out[k2] = statistics.median(v)
        out["tail_after_last_scan_ns"] = out["kernel_span"] - (out["scan_end_median"] + out["scan_end_spread"] / 2)
        bytes_ = n * a.d * (4 if a.store == "f32" else 2)
        out["gbs_by_span"] = bytes_ / out["kernel_span"]
        print(json.dumps(out))
        idx.close()


if __name__ == "__main__":
    main()
