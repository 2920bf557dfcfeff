#!/bin/bash
set -u
for i in 1 2; do
timeout 600 python bench.py --workload 10Mx768_ip_f32_k100_nq10000 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(round(j['ms_per_step'],2), round(j['value']), j['e2e'], j['roofline']['uncertified_queries_recomputed'])
"
done
