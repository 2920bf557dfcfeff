#!/bin/bash
# quick check after a host-path change: parity + error + resident + prefilter tests, then the headline bench
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_errors_gpu.py tests/test_resident_gpu.py tests/test_prefilter_gpu.py tests/test_scan_tail_gpu.py -x -q > $O/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/quick_pytest.log
timeout 600 python bench.py --no-others > $O/quick_bench.json 2> $O/quick_bench.err; echo "bench rc=$?"; tail -2 $O/quick_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/quick_bench.json'):
    if l.startswith('{'):
        j=json.loads(l); print('value', round(j['value'],2), 'e2e', round(j['e2e']['value'],2), 'ms', round(j['ms_per_step'],4), 'e2e_p50_ms', round(j['e2e']['p50_ms'],4), 'parity', j['parity']['ok'])
PY
