#!/bin/bash
set -u
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2l_pytest.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 )) s" >> $O/r2l_pytest.log
tail -5 $O/r2l_pytest.log
K="python tools/k3_probe.py 10000000 768 64 10"
timeout 300 $K > $O/r2l_plain_rows.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_rows_topk -s 2 -c 2 -f -o $O/r2l_prof_gemm_rows $K > $O/r2l_ncu.log 2>&1
tail -3 $O/r2l_ncu.log
