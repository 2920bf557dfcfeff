#!/bin/bash
set -u
O=gpurun_out
timeout 600 python tools/bench_small_batches.py > $O/r2k_small_batches_10Mx768.jsonl 2> $O/r2k.err; echo "rc=$?"; cut -c1-420 $O/r2k_small_batches_10Mx768.jsonl; tail -3 $O/r2k.err
timeout 600 python tools/bench_small_batches.py --rows 40000000 --d 384 --metric 1 > $O/r2k_small_batches_40Mx384_l2.jsonl 2>> $O/r2k.err; echo "rc=$?"; cut -c1-420 $O/r2k_small_batches_40Mx384_l2.jsonl
