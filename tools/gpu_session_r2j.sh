#!/bin/bash
set -u
O=gpurun_out
timeout 90 python tools/k3_probe.py 200000 768 64 10 > $O/r2j_probe.log 2>&1; echo "probe rc=$?"; tail -6 $O/r2j_probe.log
if grep -q "ids equal: True" $O/r2j_probe.log; then
  timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_shard_batch_gpu.py -x -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2j_pytest.log
  for nq in 2 8 64 128; do timeout 120 python tools/k3_probe.py 10000000 768 $nq 10 2>&1 | grep "second search" ; done > $O/r2j_small_batches.log 2>&1
  cat $O/r2j_small_batches.log
fi
