#!/bin/bash
set -u
O=gpurun_out
T0=$(date +%s)
timeout 900 python bench.py > $O/r2i_bench.json 2> $O/r2i_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
tail -3 $O/r2i_bench.err
