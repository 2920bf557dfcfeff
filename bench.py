#!/usr/bin/env python3
"""bench.py — the flat vector recall path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--no-others]

metric  : QPS (and p50 ms) @k=10 on a 10M x 768 fp32 flat IP index, single query per step;
          achieved HBM GB/s vs the measured B200 peak in `roofline`.
step    : one index.search(q[1,768], k=10) over the whole database (memo_cli.py:292).
value   : whole-job queries/s with the queries already resident in HBM (kernel path only): K searches enqueued
          back to back on one stream, ONE pair of CUDA events around them, max over ranks.  Consecutive launches
          overlap through programmatic dependent launch (the next scan ramps up while the previous one merges), so
          the per-launch time of the roofline is timed-region / launches; `isolated_launch_ms` is one launch alone.
e2e     : the same through the public host API — index.search(numpy q) -> numpy (D, I): pinned H2D
          of the query, the scan, D2H of the result inside the timed region.
N > 1   : the SAME database row-sharded over N GPUs (strong scaling): per-rank scan whose last CTA exchanges
          the local top-k with the peers over NVLink and merges (one kernel per GPU per search); --exchange nccl
          = one NCCL all-gather of the packed local top-k + the K4 merge kernel.
parity  : after the timed region the oracle (CPU restatement, oracle/) re-derives the last queries' answers:
          every returned (row, distance) bit-exact from the generator, order / uniqueness, and completeness on
          sampled 64k-row blocks; a mismatch fails the run (exit status 3).
other_configs : BASELINE configs 1-4 timed briefly in the same invocation (their own roofline + parity).
--impl reference : the reference's CPU implementation of the path.  faiss-cpu is not installable
          in this image, so this is the oracle port (oracle/flat_oracle.c) on ALL host cores (thread count set
          explicitly, not inherited from torchrun's OMP_NUM_THREADS=1); every step scans the full row count
          (a 1.5 GB block of generated rows streamed ceil(N/rows_in_block) times), rank 0 only.

Synthetic data: counter-based generator, database seed 1234, query seed 5678+step.
The database (30.72 GB) is far larger than the 126 MB L2, so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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
# the headline database searched with option `prefilter` (off by default): the scan kernel ranks the
# resident bf16 shadow (half the bytes), the best 40 rows are re-scored from the fp32 rows, a certificate proves the
# answer exact or the fp32 scan runs.  Reported next to the headline, never instead of it.
PREFILTER_WORKLOAD = "10Mx768_ip_f32_k10_nq1+bf16_prefilter"
WORKLOADS[PREFILTER_WORKLOAD] = WORKLOADS["10Mx768_ip_f32_k10_nq1"]
WORKLOAD_OPTIONS = {PREFILTER_WORKLOAD: {"prefilter": 1}}
DEFAULT_WORKLOAD = "10Mx768_ip_f32_k10_nq1"
# BASELINE.json configs 1-4, timed briefly next to the headline (the 10M x 768 database of config 2 is the headline's)
OTHER_CONFIGS = ["10kx384_ip_f32_k10_nq100", "1Mx768_cos_f32_k10_nq1", "10Mx768_ip_f32_k100_nq10000", "10Mx1024_cos_bf16_k10_nq1",
                 "100Mx384_l2_f32_k10_nq1"]
L2_BYTES = 126 * 1024 * 1024
DB_SEED, Q_SEED = 1234, 5678
METRIC_LABEL = {
    "10Mx768_ip_f32_k10_nq1": "QPS @k=10, 10Mx768 flat IP (single query)",
}


def metric_label(workload: str) -> str:
    return METRIC_LABEL.get(workload, f"QPS {workload}")


def workload_config(workload: str, world: int) -> dict:
    """The `config` object — identical for both arms (the driver compares them)."""
    n, d, metric, store, normalize, k, nq = WORKLOADS[workload]
    per = -(-n // world)
    return {"workload": workload, "rows": n, "d": d, "k": k, "nq": nq, "metric": "ip" if metric == 0 else "l2",
            "store": store, "normalize": normalize, "sharding": f"row-wise x{world}", "rows_per_gpu": per,
            "l2_policy": l2_policy(workload, world)}


def l2_policy(workload: str, world: int) -> str:
    n, d, metric, store, normalize, k, nq = WORKLOADS[workload]
    shard_bytes = -(-n // world) * d * (4 if store == "f32" else 2)
    if shard_bytes > 4 * L2_BYTES:
        return "database >> 126 MB L2, distinct query per step; no flush needed"
    if shard_bytes > L2_BYTES:
        return (f"database shard ({shard_bytes / 1e6:.0f} MB) > 126 MB L2 and streamed front to back every step (the front of the shard is evicted "
                "before the next step reads it again); distinct query per step; no flush")
    return (f"database shard ({shard_bytes / 1e6:.1f} MB) fits the 126 MB L2 and is not flushed between steps: a latency figure "
            "(launch-bound), no bandwidth claim; distinct queries per step")


def scan_passes(nq: int) -> int:
    """Scan launches one search issues (mirror of pick_qb in csrc/cabi.cu: query blocks of 8/4/2/1)."""
    passes, rem = 0, nq
    while rem > 0:
        qb = 8 if rem >= 8 else 4 if rem > 2 else rem
        rem -= min(qb, rem)
        passes += 1
    return passes


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            pass
    return None


def measured_tensor_peak():
    j = _peaks()
    if j and "bf16_tflops_sustained" in j:
        return float(j["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
    return 1400.0, "fallback (B200_PROFILING.md)"


def measured_peaks():
    j = _peaks()
    if j and "hbm_gbs" in j:
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha16() -> str:
    """Hash of the scan kernel's sources: profiles/ncu_traffic.json is only trusted for the build it was captured from."""
    h = hashlib.sha256()
    for name in ("scan_topk.cuh", "common.cuh"):
        h.update((ROOT / "c99_vectordb_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic(workload: str):
    """dram bytes per launch from the committed `ncu --set full` capture, or (None, why)."""
    try:
        j = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
    except Exception:
        return None, "no profiles/ncu_traffic.json"
    if j.get("kernel_source_sha16") != kernel_source_sha16():
        return None, "stale: profiles/ncu_traffic.json was captured from another build of csrc/scan_topk.cuh"
    v = j.get("workloads", {}).get(workload)
    return (v, j.get("source")) if v else (None, "no capture for this workload")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML in a
    background thread every ~4 ms (an nvidia-smi subprocess needs longer to start than an 8-GPU timed region lasts);
    the nvidia-smi query of the recipe is the fallback when the NVML binding is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []  # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.thread = None
        self.proc = None
        self.tmp = None
        self._stop = False

    def _loop(self, nv, h):
        while not self._stop:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            phys = ids[self.gpu] if self.gpu < len(ids) else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv = nv
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            nv = self.nv
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
            names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
            reasons = sorted({n for _, bits in self.samples for n, bit in names if bits & bit})
            return {"sm_mhz": statistics.median(float(c) for c, _ in self.samples), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(self.samples), "source": "NVML, sampled every ~4 ms from the first warm-up step to the end of the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
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
                "reasons": sorted(reasons), "samples": len(rows), "source": "nvidia-smi -lms 20"}


# ------------------------------------------------------------------------------------------------
# the oracle as checker and as CPU arm
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_rows(oracle, first_row, n, d, normalize, store):
    """Rows [first_row, first_row+n) exactly as the device holds them (K1 normalise, bf16 rounding)."""
    x = oracle.synth_rows(n, d, DB_SEED, first_row=first_row)
    if normalize:
        x = oracle.normalize_rows(x, oracle.ORDER_DEVICE)
    if store == "bf16":
        x = oracle.round_bf16(x)
    return x


def _better(metric, s_a, r_a, s_b, r_b):
    if s_a != s_b:
        return s_a > s_b if metric == 0 else s_a < s_b
    return r_a < r_b


def parity_check(workload: str, q_raw, D, I, blocks: int = 4, block_rows: int = 65536, seed: int = 99) -> dict:
    """Oracle check of finished searches (q_raw [m,d] as generated, D/I [m,k] host arrays, ids = global row positions):
    every returned distance recomputed bit-exact from the generator in the kernels' summation order, best-first order
    under the tie rule, unique in-range ids, and completeness over `blocks` random 64k-row blocks + the first and the
    last block: no sampled row may beat the k-th result without being in it."""
    import numpy as np

    from oracle import oracle

    oracle.set_threads(host_threads())
    n, d, metric, store, normalize, k, nq = WORKLOADS[workload]
    chunk = 8 if store == "bf16" else 4
    qn = oracle.normalize_rows(q_raw, oracle.ORDER_DEVICE) if normalize else q_raw
    bad, recomputed = [], 0
    m = q_raw.shape[0]
    for i in range(m):
        ids = I[i]
        if (ids < 0).any() or (ids >= n).any():
            bad.append(f"query {i}: id out of range")
            continue
        if len(set(ids.tolist())) != len(ids):
            bad.append(f"query {i}: duplicate ids")
        for j in range(1, k):
            if _better(metric, D[i, j], ids[j], D[i, j - 1], ids[j - 1]):
                bad.append(f"query {i}: position {j} out of order")
                break
        for j in range(k):
            row = host_rows(oracle, int(ids[j]), 1, d, normalize, store)
            s = oracle.scores(metric, row, qn[i], order=oracle.ORDER_DEVICE, chunk=chunk)[0]
            recomputed += 1
            if s != D[i, j]:
                bad.append(f"query {i} pos {j} row {int(ids[j])}: oracle {float(s)!r} vs device {float(D[i, j])!r}")
    rng = np.random.default_rng(seed)
    B = min(block_rows, n)
    offsets = [0, n - B] + [int(o) for o in rng.integers(0, max(1, n - B), size=blocks)]
    scored = 0
    for off in offsets:
        blk = host_rows(oracle, off, B, d, normalize, store)
        for i in range(m):
            s = oracle.scores(metric, blk, qn[i], order=oracle.ORDER_DEVICE, chunk=chunk)
            kth_s, kth_r = D[i, k - 1], int(I[i, k - 1])
            cand = np.nonzero(s >= kth_s)[0] if metric == 0 else np.nonzero(s <= kth_s)[0]
            have = set(I[i].tolist())
            for jrow in cand:
                r = off + int(jrow)
                if _better(metric, s[jrow], r, kth_s, kth_r) and r not in have:
                    bad.append(f"query {i}: row {r} (score {float(s[jrow])!r}) beats the k-th result but was not returned")
        scored += 1
    return {"checked": True, "ok": not bad, "queries": m, "rows_recomputed": recomputed, "blocks_scored": scored,
            "block_rows": B, "how": "oracle (oracle/flat_oracle.c, device summation order): bit-exact distances, order, "
            "uniqueness, sampled completeness", "violations": bad[:5]}


def cpu_sample_rows(n_rows: int, d: int) -> int:
    # about 1.5 GB of rows: seconds to generate, a few 100 ms per all-core scan; far larger than any host cache
    return int(min(n_rows, max(10_000, (1_500_000_000 // (d * 4)))))


def run_cpu(workload: str, steps: int, warmup: int, threads: int, step_budget_s: float = 6.0):
    """The reference path on the host: every step answers nq queries over `n` rows.  Rows come from a ~1.5 GB block of
    the generator that is streamed ceil(n / block) times per step (row positions offset per pass, lists merged), so a
    step does the full-size arithmetic and memory traffic without 30-150 GB of host memory.  If one full step would
    exceed `step_budget_s` the step covers a whole number of passes that fits and the value is scaled (flagged)."""
    import numpy as np

    from oracle import oracle

    oracle.set_threads(threads)
    n, d, metric, store, normalize, k, nq = WORKLOADS[workload]
    ns = cpu_sample_rows(n, d)
    db = oracle.synth_rows(ns, d, DB_SEED)
    if normalize:
        db = oracle.normalize_rows(db)
    if store == "bf16":
        db = oracle.round_bf16(db)
    rowpar = nq < threads  # few queries: split the rows over the threads; many: one thread per query (faiss's way)
    full_passes = -(-n // ns)

    def one_pass(q, p, rows):
        Dp, Ip = oracle.search(metric, db[:rows], q, k, rowpar=rowpar)
        Ip[Ip >= 0] += p * ns
        return Dp, Ip

    def step(q, passes):
        parts = []
        for p in range(passes):
            rows = min(ns, n - p * ns)
            parts.append(one_pass(q, p, rows))
        if len(parts) == 1:
            return parts[0]
        return oracle.merge_topk(metric, np.stack([x[0] for x in parts]), np.stack([x[1] for x in parts]))

    # calibrate on one pass
    q = oracle.synth_rows(nq, d, Q_SEED - 1)
    if normalize:
        q = oracle.normalize_rows(q)
    t0 = time.perf_counter()
    one_pass(q, 0, ns)
    pass_s = time.perf_counter() - t0
    passes = full_passes if pass_s * full_passes <= step_budget_s else max(1, int(step_budget_s / pass_s))
    rows_per_step = n if passes == full_passes else passes * ns
    times = []
    for s in range(warmup + steps):
        q = oracle.synth_rows(nq, d, Q_SEED + s)
        if normalize:
            q = oracle.normalize_rows(q)
        t0 = time.perf_counter()
        step(q, passes)
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    scale = rows_per_step / n
    qps = nq * len(times) / total * scale
    return {
        "value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "extrapolated": passes != full_passes,
        "sample": f"{len(times)} steps x {nq} quer{'y' if nq == 1 else 'ies'}, each over {rows_per_step} of {n} rows "
                  f"({passes} pass{'es' if passes != 1 else ''} over a {ns}-row block of the generator, {ns * d * 4 / 1e9:.2f} GB"
                  + ("" if passes == full_passes else f"; value scaled by {rows_per_step}/{n}") + "); "
                  f"{threads} threads, {'rows split over the threads (our extension)' if rowpar else 'one thread per query as faiss runs batches'}; "
                  "oracle port of faiss flat search (faiss-cpu not installable here)",
        "ms_per_step": 1e3 * total / len(times) / scale,
        "p50_ms": 1e3 * statistics.median(times) / scale,
        "timed_s": total,
    }


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    cb = run_cpu(a.workload, a.steps, a.warmup, threads)
    n, d, metric, store, normalize, k, nq = WORKLOADS[a.workload]
    line = {
        "impl": "reference", "metric": metric_label(a.workload), "value": cb["value"], "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_step"], "p50_ms": cb["p50_ms"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a.workload, a.gpus),
        "cpu_baseline": {k2: cb[k2] for k2 in ("value", "unit", "cores", "kind", "sample", "extrapolated")},
        "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU port of the reference path (the whole database on one host, whatever --gpus says)",
    }
    if a.one_core:
        c1 = run_cpu(a.workload, max(1, min(a.steps, 3)), 1, 1, step_budget_s=4.0)
        line["one_core_as_faiss"] = {k2: c1[k2] for k2 in ("value", "unit", "cores", "sample", "extrapolated")}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, a):
        import torch
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus and self.world > 1:
            raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.torch, self.dist = torch, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def min_over_ranks(self, value: int) -> int:
        t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())


def run_workload(env: Env, a, workload: str, steps: int, warmup: int, *, headline: bool, check_blocks: int = 4) -> dict:
    """Build the (sharded) index of `workload`, time `steps` searches (device-resident and through the host API),
    check the last answers with the oracle, return the result line (rank 0) — every rank takes part."""
    import numpy as np

    import c99_vectordb_b200 as m
    from c99_vectordb_b200 import _cabi
    from c99_vectordb_b200.sharded import ShardedIndexFlat, shard_range

    torch, dist = env.torch, env.dist
    world, rank, dev = env.world, env.rank, env.dev
    n, d, metric, store, normalize, k, nq = WORKLOADS[workload]
    idx = ShardedIndexFlat(d, metric, store=store, normalize=normalize)
    base = idx.local.index
    fused = world > 1 and a.exchange == "fused" and idx.enable_fused_exchange()
    base.set_option("queries_stable", 1)  # every query of the timed loops is resident and final before the loop starts
    for name, val in (a.option or []):
        base.set_option(name, int(val))
    variant_opts = WORKLOAD_OPTIONS.get(workload, {})
    for name, val in variant_opts.items():
        base.set_option(name, int(val))
    prefilter = bool(variant_opts.get("prefilter")) and world == 1
    t0 = time.time()
    ok_build = 1
    try:
        idx.add_synthetic(n, DB_SEED)
    except RuntimeError as e:  # e.g. 153.6 GB next to other allocations: every rank must agree to skip
        ok_build, build_err = 0, str(e)
    if env.min_over_ranks(ok_build) == 0:
        idx.local.index.close()
        return {"workload": workload, "skipped": build_err if not ok_build else "another rank could not build its shard"}
    torch.cuda.synchronize()
    build_s = time.time() - t0
    lo, hi = shard_range(n, world, rank)
    elem = 4 if store == "f32" else 2
    if prefilter:
        elem = 2  # the scan streams the bf16 shadow; the fp32 rows are touched only by the re-rank (40 rows per query)
    bytes_per_scan_total = n * d * elem  # algorithmic bytes of one pass over the whole database

    # queries: device resident for `value`, host for `e2e` (distinct per step)
    total_steps = warmup + steps
    q_all = torch.empty((total_steps, nq, d), dtype=torch.float32, device=dev)
    for s in range(total_steps):
        _cabi.check(_cabi.load().b200_synth_rows_dev(q_all[s].data_ptr(), nq, d, Q_SEED + s, 0, 0, C.c_void_p(1)))
    torch.cuda.synchronize()
    q_host = q_all.cpu().numpy()
    single_kernel = (world == 1 or fused)

    # ---- value: device-resident queries, one CUDA-event pair around K back-to-back searches ----
    sampler = ClockSampler(env.local_rank) if headline else None
    if sampler and rank == 0:
        sampler.start()
    if nq > 1 and base.get_option("gemm_min_nq") > 0:
        # warm-up of the batched path's RARE branches too: starved thresholds make certificates fail, so the widened
        # second pass and the exact-scan fallback run once before anything is timed (CUDA loads a kernel on first use;
        # a query set that needs a retry would otherwise pay hundreds of ms of one-time loading inside a timed step)
        keep = base.get_option("gemm_emit_factor")
        base.set_option("gemm_emit_factor", 2)
        idx.search_device(q_all[0][: min(nq, 512)].contiguous(), k)
        base.set_option("gemm_emit_factor", keep)
    for s in range(warmup):
        idx.search_device(q_all[s], k)
    env.barrier()
    if sampler and rank == 0:
        sampler.samples.clear()  # keep what is sampled under the timed load only
    launches0 = idx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = []
    ev0.record()
    for s in range(steps):
        q = q_all[warmup + s]
        if single_kernel or nq > 1:  # batches: K3 per shard + one all-gather + merge/certificate (sharded.py)
            idx.search_device(q, k)
        else:
            # NCCL exchange: time the scan launches alone for the roofline, the whole step for the metric
            mine, gathered, D, I, nbytes, off_d = idx._buffers(nq, k)
            I_loc = mine[: nq * k * 8].view(torch.int64).view(nq, k)
            D_loc = mine[off_d: off_d + nq * k * 4].view(torch.float32).view(nq, k)
            kev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            kev[-1][0].record()
            idx._local_search(q, k, D_loc, I_loc)
            kev[-1][1].record()
            dist.all_gather_into_tensor(gathered, mine)
            idx._merge(gathered, nq, k, nbytes, off_d, D, I)
    ev1.record()
    env.barrier()
    clocks = sampler.stop() if (sampler and rank == 0) else None
    launches = idx.launch_count - launches0
    total_ms = ev0.elapsed_time(ev1)
    gemm_used = bool(base.get_option("stat_gemm_used"))

    # ---- isolated launches: per-search device latency (events around each search, no overlap with neighbours) ----
    iso = []
    for s in range(min(steps, 20)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        e0.record()
        idx.search_device(q_all[warmup + s], k)
        e1.record()
        torch.cuda.synchronize()
        iso.append(e0.elapsed_time(e1))
    scan_ms = [x.elapsed_time(y) for x, y in kev] if kev else [total_ms / steps]
    total_ms, iso_p50, scan_avg_ms = env.max_over_ranks([total_ms, statistics.median(iso), sum(scan_ms) / len(scan_ms)])
    value = steps * nq / (total_ms * 1e-3)

    # ---- e2e: host query -> host result through the public API ----
    # index.search(numpy) -> numpy: the C ABI host entry on one GPU, the sharded host entry otherwise
    host_search = idx.local.search if world == 1 else idx.search
    for s in range(min(warmup, 3)):
        host_search(q_host[s], k)  # the SAME call as the timed one (its pinned staging is sized on first use)
    env.barrier()
    e2e_lat = []
    t_start = time.perf_counter()
    for s in range(steps):
        t1 = time.perf_counter()
        D_h, I_h = host_search(q_host[warmup + s], k)
        e2e_lat.append(time.perf_counter() - t1)
    env.barrier()
    e2e_total, = env.max_over_ranks([time.perf_counter() - t_start])
    e2e_qps = steps * nq / e2e_total
    # per rank: a sharded host batch uploads 1/world of the queries and all-gathers the rest over NVLink (sharded.py)
    h2d_bytes = nq * d * 4
    if world > 1 and h2d_bytes >= idx.query_allgather_min_bytes:
        h2d_bytes = -(-nq // world) * d * 4

    # ---- parity: the oracle re-derives the last answers of the device-resident path (and the host path agrees) ----
    chk_steps = [total_steps - 1 - j for j in range(min(3, steps))] if nq == 1 else [total_steps - 1]
    res = []
    for s in chk_steps:  # every rank searches (the exchange is collective); rank 0 checks
        Dd, Id = idx.search_device(q_all[s], k)
        torch.cuda.synchronize()
        res.append((Dd.cpu().numpy().copy(), Id.cpu().numpy().copy()))
    idx.check_exchange()
    host_agrees = bool((res[0][1] == I_h).all() and (res[0][0] == D_h).all())
    ranks_agree = True
    if world > 1:
        mine_I = torch.from_numpy(np.concatenate([r[1].reshape(-1) for r in res])).to(dev)
        ref_I = mine_I.clone()
        dist.broadcast(ref_I, 0)
        ranks_agree = bool(env.min_over_ranks(int(bool((ref_I == mine_I).all().item()))))
    parity = None
    if rank == 0 and not a.no_parity:
        if nq == 1:
            qs = np.stack([q_host[s][0] for s in chk_steps])
            Ds, Is = np.stack([r[0][0] for r in res]), np.stack([r[1][0] for r in res])
        else:
            m_chk = min(3, nq)
            qs, Ds, Is = q_host[chk_steps[0]][-m_chk:], res[0][0][-m_chk:], res[0][1][-m_chk:]
        parity = parity_check(workload, qs, Ds, Is, blocks=check_blocks)
        parity["host_api_agrees_with_device_api"] = host_agrees
        parity["ranks_agree"] = ranks_agree
        parity["ok"] = bool(parity["ok"] and host_agrees and ranks_agree)

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        launch_bytes = (hi - lo) * d * elem  # algorithmic bytes one scan launch streams on this rank
        scans_per_step = scan_passes(nq)  # nq > 8 -> several passes over the database per step
        launch_ms = scan_avg_ms / scans_per_step
        achieved = launch_bytes / (launch_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(workload) if world == 1 else (None, "captured on one GPU only")
        line = {
            "metric": metric_label(workload),
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "p50_ms": iso_p50, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if store == "f32" else "bf16-stored/f32-accumulate", "data": "synthetic",
            "config": workload_config(workload, world),
            "run": {"build_s": round(build_s, 3),
                    "exchange": "none" if world == 1 else (
                        "fused: last CTA posts the local top-k into every peer's buffer over NVLink (CUDA IPC) as self-validating "
                        "words, polls its own buffer, merges — one kernel per GPU" if fused else "NCCL all_gather of packed (I,D)[nq,k] + K4 merge kernel"),
                    "overlap": "programmatic dependent launch, queries_stable=1" if base.get_option("scan_pdl") else "none"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "scan_topk_kernel", "bytes_per_launch": launch_bytes,
                         "scan_launches_per_step": scans_per_step,
                         "note": "avg_launch_ms = CUDA-event time of the timed region / scan launches in it (back-to-back launches "
                                 "overlap their tails by design); isolated_launch_ms = events around one search alone",
                         "avg_launch_ms": launch_ms, "isolated_launch_ms": iso_p50 / scans_per_step, "peak_source": peak_src,
                         "whole_job_gbs": bytes_per_scan_total * scans_per_step / (total_ms / steps * 1e-3) / 1e9},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": nq * k * 12,
                    "p50_ms": 1e3 * statistics.median(e2e_lat), "max_ms": 1e3 * max(e2e_lat),
                    "timing": "host wall clock around index.search(numpy)"},
            "gpu_launches": int(launches),
            "parity": parity,
        }
        if clocks is not None:
            line["clocks"] = clocks
        if prefilter:
            line["roofline"]["kernel"] = "scan_topk_kernel over the resident bf16 shadow + rerank_kernel (exact fp32 re-score of 40 rows, certificate)"
            line["roofline"]["traffic"], line["roofline"]["traffic_source"] = None, "not captured for this variant"
            line["roofline"]["bytes_note"] = ("bytes_per_launch = rows x d x 2: what this variant has to read (the bf16 shadow); against the "
                                              "fp32 rows' 30.72 GB the same time corresponds to %.0f GB/s" % (n * d * 4 / (launch_ms * 1e-3) / 1e9))
            line["prefilter"] = {"last_search_used_it": bool(base.get_option("stat_prefilter_used")),
                                 "uncertified_searches_recomputed_by_the_fp32_scan": int(base.get_option("stat_prefilter_fallbacks")),
                                 "note": "option prefilter=1 (off by default); results are proven exact by the certificate or recomputed"}
    if store == "bf16":
        # bf16 storage is lossy: report recall@k against the same rows stored in fp32 (north_star)
        ref = ShardedIndexFlat(d, metric, store="f32", normalize=normalize)
        ok_ref = 1
        try:
            ref.add_synthetic(n, DB_SEED)
        except RuntimeError:
            ok_ref = 0
        if env.min_over_ranks(ok_ref):
            hits = tot = 0
            for s in range(min(total_steps, 50)):
                _, I16 = idx.search_device(q_all[s], k)
                a16 = I16.cpu().numpy()
                _, I32 = ref.search_device(q_all[s], k)
                a32 = I32.cpu().numpy()
                for r in range(nq):
                    hits += len(set(a16[r].tolist()) & set(a32[r].tolist()))
                    tot += k
            if line is not None:
                line[f"recall_at_{k}_vs_fp32_rows"] = hits / tot
        ref.local.index.close()
    if gemm_used and line is not None:
        # batched path: the dominant kernel is the tcgen05 emit pass; algorithmic flops = 2 nq N d
        gstat = (lambda name: base.get_option("stat_gemm_" + name)) if world == 1 or not idx.last_batch_stats else (lambda name: idx.last_batch_stats[name])
        p2_ms = gstat("pass2_us") / 1e3
        flops = 2.0 * nq * (hi - lo) * d
        tpeak, tsrc = measured_tensor_peak()
        line["roofline"] = {"bound": "tensor", "achieved": flops / (p2_ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                            "frac": flops / (p2_ms * 1e-3) / 1e12 / tpeak, "traffic": None, "kernel": "gemm_topk_kernel (emit pass)",
                            "flops_per_launch": flops, "avg_launch_ms": p2_ms, "peak_source": tsrc,
                            "whole_step_tflops": 2.0 * nq * n * d / (total_ms / steps * 1e-3) / 1e12,
                            "pass1_ms": gstat("pass1_us") / 1e3,
                            "rerank_ms": gstat("rerank_us") / 1e3,
                            "uncertified_queries_recomputed": (base.get_option("stat_gemm_fallbacks") if world == 1
                                                               else idx.last_batch_uncertified),
                            "candidates_per_query": (base.get_option("stat_gemm_cand_total") / nq if world == 1 else None),
                            "note": "bf16 tensor-core pass (tcgen05, TMEM accumulators) + exact fp32 re-rank; "
                                    "launch time is the kernel's own CUDA-event bracket from the last step"}
    idx.local.index.close()
    del idx
    torch.cuda.empty_cache()
    return line


def main_b200(a):
    env = Env(a)
    line = run_workload(env, a, a.workload, a.steps, a.warmup, headline=True)
    others = []
    if not a.no_others and a.workload == DEFAULT_WORKLOAD:
        for w in OTHER_CONFIGS + ([PREFILTER_WORKLOAD] if env.world == 1 else []):
            if env.world > 1 and WORKLOADS[w][0] < 1_000_000:
                continue  # config 0 (10k rows) is the CPU-runnable case: one GPU; sharding 10k rows measures nothing
            nq = WORKLOADS[w][6]
            st = max(5, min(a.steps, 10)) if nq == 1 else 5
            try:
                r = run_workload(env, a, w, st, 3, headline=False, check_blocks=2)
            except Exception as e:  # an extra line must never take the headline down; say what happened
                r = {"workload": w, "error": f"{type(e).__name__}: {e}"} if env.rank == 0 else None
                if env.world > 1:
                    raise
            if env.rank == 0 and r is not None:
                if "skipped" in r or "error" in r:
                    others.append(r)
                else:
                    keep = {k2: r[k2] for k2 in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "p50_ms", "dtype",
                                                 "roofline", "e2e", "gpu_launches", "parity", "prefilter") if k2 in r}
                    keep["workload"] = w
                    keep["l2_policy"] = l2_policy(w, env.world)
                    if w == "10kx384_ip_f32_k10_nq100" and not a.no_cpu:
                        # BASELINE config 0 is the reference's own CPU-runnable case: the host does it in full
                        cb0 = run_cpu(w, 5, 1, host_threads())
                        keep["cpu_baseline"] = {k2: cb0[k2] for k2 in ("value", "unit", "cores", "kind", "sample", "extrapolated")}
                    for k2 in r:
                        if k2.startswith("recall_at_"):
                            keep[k2] = r[k2]
                    others.append(keep)
    rc = 0
    if env.rank == 0:
        line["other_configs"] = others
        if not a.no_cpu and env.world == 1:
            cbN = run_cpu(a.workload, 5, 1, host_threads())
            line["cpu_baseline"] = {k2: cbN[k2] for k2 in ("value", "unit", "cores", "kind", "sample", "extrapolated")}
        print(json.dumps(line))
        checks = [line.get("parity")] + [o.get("parity") for o in others]
        if any(c is not None and not c["ok"] for c in checks):
            print("bench.py: PARITY FAILURE (see the `parity` objects of the line above)", file=sys.stderr)
            rc = 3
    if env.world > 1:
        env.dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--option", nargs=2, action="append", metavar=("NAME", "VALUE"), help="native tuning option")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the last answers")
    ap.add_argument("--one-core", action="store_true", help="reference arm: also time one thread (how faiss runs a single query)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="multi-GPU top-k exchange")
    a = ap.parse_args()
    if a.impl == "reference":
        a.warmup = max(a.warmup, 1)
        return main_reference(a)
    a.warmup = max(a.warmup, 3)
    return main_b200(a)


if __name__ == "__main__":
    sys.exit(main())
