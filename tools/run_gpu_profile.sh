#!/bin/bash
# one GPU box visit: tests, bench, K3 probe, launch list, ncu --set full of K2 and K3 (each after its plain run exited 0)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/prof_pytest.log 2>&1; echo "pytest rc=$?" >> $O/prof_pytest.log
tail -3 $O/prof_pytest.log
timeout 600 python bench.py > $O/prof_bench.json 2> $O/prof_bench.err; echo "bench rc=$?"
timeout 300 python tools/k3_probe.py 10000000 768 10000 100 > $O/prof_k3probe.log 2>&1; tail -4 $O/prof_k3probe.log
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-others --no-parity"
timeout 300 $B > $O/prof_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/prof_launches.csv $B > $O/prof_ncu1.log 2>&1
timeout 300 $B > $O/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 2 -f -o $O/prof_prof_scan $B > $O/prof_ncu2.log 2>&1
K="python tools/k3_probe.py 2000000 768 4096 100"
timeout 300 $K > $O/prof_plain_k3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 2 -f -o $O/prof_prof_gemm $K > $O/prof_ncu3.log 2>&1
ls -la $O | tail -12
