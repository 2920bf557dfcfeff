#!/bin/bash
# final single-GPU evidence: full GPU test suite, smoke, the default bench line (+ reference arm), launch list
set -u
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 )) s" >> $O/final_pytest.log
tail -4 $O/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/final_smoke.log
T0=$(date +%s)
timeout 900 python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
T0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 --one-core > $O/final_bench_ref.json 2> $O/final_bench_ref.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 )) s"
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-others --no-parity"
timeout 300 $B > $O/final_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv $B > $O/final_ncu1.log 2>&1
echo "launch list rc=$?"
