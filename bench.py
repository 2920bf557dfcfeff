#!/usr/bin/env python3
"""bench.py — the flat vector recall path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

metric  : QPS (and p50 ms) @k=10 on a 10M x 768 fp32 flat IP index, single query per step;
          achieved HBM GB/s vs the measured B200 peak in `roofline`.
step    : one index.search(q[1,768], k=10) over the whole database (memo_cli.py:292).
value   : whole-job queries/s with the query already resident in HBM (kernel path only),
          CUDA-event timed on the launching stream, max over ranks.
e2e     : the same through the public host API — index.search(numpy q) -> numpy (D, I): pinned H2D
          of the query, the scan, D2H of the result inside the timed region.
N > 1   : the SAME database row-sharded over N GPUs (strong scaling): per-rank scan, one NCCL
          all-gather of the packed local top-k, K4 merge kernel on every rank.
--impl reference : the reference's CPU implementation of the path.  faiss-cpu is not installable
          in this image, so this is the oracle port (oracle/flat_oracle.c), all host threads,
          timed on a bounded sample of the same workload (rank 0 only).

Synthetic data: counter-based generator, database seed 1234, query seed 5678+step.
The database (30.72 GB) is far larger than the 126 MB L2, so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (rows, d, metric, store, normalize, k, nq)
    "10Mx768_ip_f32_k10_nq1": (10_000_000, 768, 0, "f32", False, 10, 1),
    "1Mx768_cos_f32_k10_nq1": (1_000_000, 768, 0, "f32", True, 10, 1),
    "100Mx384_l2_f32_k10_nq1": (100_000_000, 384, 1, "f32", False, 10, 1),
    "10Mx1024_cos_bf16_k10_nq1": (10_000_000, 1024, 0, "bf16", True, 10, 1),
    "10kx384_ip_f32_k10_nq100": (10_000, 384, 0, "f32", True, 10, 100),
    "10Mx768_ip_f32_k100_nq10000": (10_000_000, 768, 0, "f32", False, 100, 10_000),  # config 2: tcgen05 batched path
}
DEFAULT_WORKLOAD = "10Mx768_ip_f32_k10_nq1"
DB_SEED, Q_SEED = 1234, 5678


def scan_passes(nq: int) -> int:
    """Scan launches one search issues (mirror of pick_qb in csrc/cabi.cu: query blocks of 8/4/2/1)."""
    passes, rem = 0, nq
    while rem > 0:
        qb = 8 if rem >= 8 else 4 if rem > 2 else rem
        rem -= min(qb, rem)
        passes += 1
    return passes


def measured_tensor_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 1400.0, "fallback (B200_PROFILING.md)"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            j = json.loads(p.read_text())
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.tmp.flush()
        rows = []
        try:
            for line in open(self.tmp.name):
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 9:
                    rows.append(f)
        finally:
            try:
                os.unlink(self.tmp.name)
            except OSError:
                pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU implementation of the path (oracle port), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample_rows(n_rows: int, d: int) -> int:
    # about 1.5 GB of rows: seconds to generate, a few 100 ms per all-core scan
    return int(min(n_rows, max(10_000, (1_500_000_000 // (d * 4)))))


def run_cpu(workload: str, steps: int, warmup: int, rowpar: bool):
    from oracle import oracle

    n, d, metric, store, normalize, k, nq = WORKLOADS[workload]
    ns = cpu_sample_rows(n, d)
    db = oracle.synth_rows(ns, d, DB_SEED)
    if normalize:
        db = oracle.normalize_rows(db)
    if store == "bf16":
        db = oracle.round_bf16(db)
    threads = oracle.max_threads() if (rowpar or nq > 1) else 1
    times = []
    for s in range(warmup + steps):
        q = oracle.synth_rows(nq, d, Q_SEED + s)
        if normalize:
            q = oracle.normalize_rows(q)
        t0 = time.perf_counter()
        oracle.search(metric, db, q, k, rowpar=rowpar)
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    qps_sample = nq * len(times) / total
    qps_full = qps_sample * ns / n  # a full-size scan reads n/ns times the sample's bytes
    return {
        "value": qps_full, "unit": "queries/s", "cores": threads, "kind": "port",
        "sample": f"{len(times)} steps x {nq} quer{'y' if nq == 1 else 'ies'} over the first {ns} of {n} rows "
                  f"({ns * d * 4 / 1e9:.2f} GB, same generator); QPS scaled by {ns}/{n}; "
                  f"{'rows split over all threads (our extension)' if rowpar else 'one thread per query as faiss runs nq<20'}; "
                  "oracle port of faiss flat search (faiss-cpu not installable here)",
        "ms_per_step_sample": 1e3 * total / len(times),
        "p50_ms_full_est": 1e3 * statistics.median(times) * n / ns,
    }


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(a.steps, 20)
    cb = run_cpu(a.workload, steps, min(a.warmup, 2), rowpar=True)
    n, d, metric, store, normalize, k, nq = WORKLOADS[a.workload]
    line = {
        "impl": "reference", "metric": "QPS @k=10, 10Mx768 flat IP (single query)", "value": cb["value"], "unit": "queries/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": min(a.warmup, 2), "ms_per_step": 1e3 / cb["value"] * nq,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": a.workload, "rows": n, "d": d, "k": k, "nq": nq,
                   "note": "CPU port of the reference path on a bounded sample, scaled to the full database"},
        "cpu_baseline": {k2: cb[k2] for k2 in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import c99_vectordb_b200 as m
    from c99_vectordb_b200 import _cabi
    from c99_vectordb_b200.sharded import ShardedIndexFlat, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, d, metric, store, normalize, k, nq = WORKLOADS[a.workload]
    idx = ShardedIndexFlat(d, metric, store=store, normalize=normalize)
    base = idx.local.index
    fused = world > 1 and a.exchange == "fused" and idx.enable_fused_exchange()
    for name, val in (a.option or []):
        base.set_option(name, int(val))
    t0 = time.time()
    idx.add_synthetic(n, DB_SEED)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    lo, hi = shard_range(n, world, rank)
    elem = 4 if store == "f32" else 2
    bytes_per_scan_total = n * d * elem  # algorithmic bytes of one pass over the whole database

    # queries: device resident for `value`, host for `e2e` (distinct per step)
    total_steps = a.warmup + a.steps
    q_all = torch.empty((total_steps, nq, d), dtype=torch.float32, device=dev)
    for s in range(total_steps):
        _cabi.check(_cabi.load().b200_synth_rows_dev(q_all[s].data_ptr(), nq, d, Q_SEED + s, 0, 0, C.c_void_p(1)))
    torch.cuda.synchronize()
    q_host = q_all.cpu().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident queries, CUDA events on the launching stream ----
    for s in range(a.warmup):
        idx.search_device(q_all[s], k)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = idx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    ev[0].record()
    for s in range(a.steps):
        q = q_all[a.warmup + s]
        if world == 1 or fused:
            kev[s][0].record()
            idx.search_device(q, k)
            kev[s][1].record()
        else:
            # time the scan launches alone for the roofline, the whole step for the metric
            mine, gathered, D, I, nbytes, off_d = idx._buffers(nq, k)
            I_loc = mine[: nq * k * 8].view(torch.int64).view(nq, k)
            D_loc = mine[off_d: off_d + nq * k * 4].view(torch.float32).view(nq, k)
            kev[s][0].record()
            idx._local_search(q, k, D_loc, I_loc)
            kev[s][1].record()
            dist.all_gather_into_tensor(gathered, mine)
            idx._merge(gathered, nq, k, nbytes, off_d, D, I)
        ev[s + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = idx.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    scan_ms = [x.elapsed_time(y) for x, y in kev]
    t = torch.tensor([total_ms, statistics.median(step_ms), sum(scan_ms) / len(scan_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, p50_ms, scan_avg_ms = (float(x) for x in t.tolist())
    value = a.steps * nq / (total_ms * 1e-3)

    # ---- e2e: host query -> host result through the public API ----
    for s in range(min(a.warmup, 3)):
        idx.search(q_host[s], k)
    barrier()
    e2e_lat = []
    t_start = time.perf_counter()
    for s in range(a.steps):
        t1 = time.perf_counter()
        if world == 1:
            D_h, I_h = idx.local.search(q_host[a.warmup + s], k)  # index.search(numpy) -> numpy (C ABI host entry)
        else:
            D_h, I_h = idx.search(q_host[a.warmup + s], k)
        e2e_lat.append(time.perf_counter() - t1)
    barrier()
    e2e_total = time.perf_counter() - t_start
    t = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_total = float(t.item())
    e2e_qps = a.steps * nq / e2e_total

    # ---- sanity: the last e2e answer equals the device-resident answer ----
    Dd, Id = idx.search_device(q_all[total_steps - 1], k)
    torch.cuda.synchronize()
    ok = bool((Id.cpu().numpy() == I_h).all())

    if rank == 0:
        peak, peak_src = measured_peaks()
        gemm_used = bool(base.get_option("stat_gemm_used"))
        launch_bytes = (hi - lo) * d * elem  # algorithmic bytes one scan launch streams on this rank
        scans_per_step = scan_passes(nq)  # nq > 8 -> several passes over the database per step
        scan_avg_ms = scan_avg_ms / scans_per_step
        achieved = launch_bytes / (scan_avg_ms * 1e-3) / 1e9
        line = {
            "metric": "QPS @k=10, 10Mx768 flat IP (single query)" if a.workload == DEFAULT_WORKLOAD else f"QPS {a.workload}",
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total_ms / a.steps, "p50_ms": p50_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if store == "f32" else "bf16-stored/f32-accumulate", "data": "synthetic",
            "config": {"workload": a.workload, "rows": n, "d": d, "k": k, "nq": nq, "metric": "ip" if metric == 0 else "l2",
                       "sharding": f"row-wise x{world}", "rows_per_gpu": hi - lo,
                       "l2_policy": "database >> 126 MB L2, distinct query per step; no flush needed",
                       "build_s": round(build_s, 3), "ids_consistent_host_vs_device": ok,
                       "exchange": "none" if world == 1 else (
                           "fused: last CTA stores the local top-k into every peer's buffer over NVLink (CUDA IPC), flags, waits, merges"
                           " — one kernel per GPU" if fused else "NCCL all_gather of packed (I,D)[nq,k] + K4 merge kernel")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "scan_topk_kernel", "bytes_per_launch": launch_bytes, "scan_launches_per_step": scans_per_step,
                         "note": "avg_launch_ms = CUDA-event time around the search on its stream / scan launches"
                                 + (" (includes the ~2 us query-normalise kernel)" if normalize else ""),
                         "avg_launch_ms": scan_avg_ms, "peak_source": peak_src,
                         "whole_job_gbs": bytes_per_scan_total * scans_per_step / (total_ms / a.steps * 1e-3) / 1e9},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12,
                    "p50_ms": 1e3 * statistics.median(e2e_lat), "timing": "host wall clock around index.search(numpy)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        try:
            tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get(a.workload)
            if tr and world == 1:
                line["roofline"]["traffic"] = tr
        except Exception:
            pass
        if store == "bf16" and world == 1:
            # bf16 storage is lossy: report recall@k against the same rows stored in fp32 (north_star)
            ref = m.IndexFlat(d, metric, normalize=normalize)
            ref.add_synthetic(n, DB_SEED)
            hits = tot = 0
            for s in range(min(total_steps, 50)):
                _, I16 = idx.search_device(q_all[s], k)
                _, I32 = ref.search_device(q_all[s], k)
                torch.cuda.synchronize()
                a16, a32 = I16.cpu().numpy(), I32.cpu().numpy()
                for r in range(nq):
                    hits += len(set(a16[r].tolist()) & set(a32[r].tolist()))
                    tot += k
            line["config"][f"recall_at_{k}_vs_fp32_rows"] = hits / tot
            ref.close()
        if gemm_used:
            # batched path: the dominant kernel is the tcgen05 emit pass; algorithmic flops = 2 nq N d
            p2_ms = base.get_option("stat_gemm_pass2_us") / 1e3
            flops = 2.0 * nq * (hi - lo) * d
            tpeak, tsrc = measured_tensor_peak()
            line["roofline"] = {"bound": "tensor", "achieved": flops / (p2_ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                                "frac": flops / (p2_ms * 1e-3) / 1e12 / tpeak, "traffic": None, "kernel": "gemm_topk_kernel (emit pass)",
                                "flops_per_launch": flops, "avg_launch_ms": p2_ms, "peak_source": tsrc,
                                "pass1_ms": base.get_option("stat_gemm_pass1_us") / 1e3,
                                "rerank_ms": base.get_option("stat_gemm_rerank_us") / 1e3,
                                "uncertified_queries_recomputed": base.get_option("stat_gemm_fallbacks"),
                                "candidates_per_query": base.get_option("stat_gemm_cand_total") / nq,
                                "note": "bf16 tensor-core pass (tcgen05, TMEM accumulators) + exact fp32 re-rank; "
                                        "launch time is the kernel's own CUDA-event bracket from the last step"}
        if world == 1 and not a.no_cpu:
            cb1 = run_cpu(a.workload, 3, 1, rowpar=False)
            cbN = run_cpu(a.workload, 8, 1, rowpar=True)
            line["cpu_baseline"] = {k2: cbN[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["one_core_as_faiss"] = {"value": cb1["value"], "cores": 1, "sample": cb1["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--option", nargs=2, action="append", metavar=("NAME", "VALUE"), help="native tuning option")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="multi-GPU top-k exchange")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        return main_reference(a)
    return main_b200(a)


if __name__ == "__main__":
    sys.exit(main())
