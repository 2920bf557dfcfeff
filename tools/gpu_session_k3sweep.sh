#!/bin/bash
set -u
O=gpurun_out
: > $O/k3sweep.jsonl
for ct in 0 8 16 24 48 96; do
  timeout 300 python bench.py --workload 10Mx768_ip_f32_k100_nq10000 --steps 4 --warmup 3 --no-cpu --no-parity --option gemm_chunk_tiles $ct 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print(json.dumps({'chunk_tiles': $ct, 'ms_per_step': round(j['ms_per_step'],2), 'emit_ms': r['avg_launch_ms'], 'frac': round(r['frac'],4), 'pass1_ms': r['pass1_ms'], 'rerank_ms': r['rerank_ms'], 'uncert': r['uncertified_queries_recomputed']}))
" >> $O/k3sweep.jsonl
done
cat $O/k3sweep.jsonl
