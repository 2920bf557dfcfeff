"""First-light probe of the K3 path on one GPU: small case, prints stats; run under `timeout`."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import c99_vectordb_b200 as m
from oracle import oracle

n, d, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
idx = m.IndexFlat(d, 0)
idx.set_option("gemm_min_nq", 32)
if len(sys.argv) > 5: idx.set_option("gemm_cta_group", int(sys.argv[5]))
if len(sys.argv) > 6: idx.set_option("gemm_min_rows", int(sys.argv[6]))
idx.add_synthetic(n, 1234)
q = oracle.synth_rows(nq, d, 5678)
t0 = time.time(); D, I = idx.search(q, k); t1 = time.time()
print("first search s", round(t1 - t0, 3), {s: idx.get_option(s) for s in ("stat_gemm_used", "stat_gemm_fallbacks", "stat_gemm_cand_total", "stat_gemm_pass1_us", "stat_gemm_pass2_us", "stat_gemm_rerank_us")}, flush=True)
ts=[]
for _ in range(20):
    t0 = time.time(); D, I = idx.search(q, k); ts.append(time.time()-t0)
t0=0; t1=sorted(ts)[10]
print("second search s", round(t1 - t0, 3), {s: idx.get_option(s) for s in ("stat_gemm_fallbacks", "stat_gemm_cand_total", "stat_gemm_pass1_us", "stat_gemm_pass2_us", "stat_gemm_rerank_us")}, flush=True)
if n <= 2_000_000:
    db = oracle.synth_rows(n, d, 1234)
    m_ = min(nq, 64)
    Dw, Iw = oracle.search(0, db, q[:m_], k, order=oracle.ORDER_DEVICE)
    print("ids equal:", bool((I[:m_] == Iw).all()), "dist equal:", bool((D[:m_] == Dw).all()), flush=True)
    if not (I[:m_] == Iw).all():
        bad = np.nonzero((I[:m_] != Iw).any(axis=1))[0]
        print("bad queries", bad[:10], I[bad[0]][:10], Iw[bad[0]][:10])
flops = 2.0 * nq * n * d
p2 = idx.get_option("stat_gemm_pass2_us")
if p2: print("pass2 TFLOP/s", round(flops / (p2 * 1e-6) / 1e12, 1))
