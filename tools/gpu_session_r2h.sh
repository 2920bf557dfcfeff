#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_prefilter_gpu.py tests/test_gemm_gpu.py -x -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2h_pytest.log
tail -30 $O/r2h_pytest.log
timeout 600 python tools/bench_prefilter.py > $O/r2h_prefilter.jsonl 2> $O/r2h_prefilter.err; echo "rc=$?"; cat $O/r2h_prefilter.jsonl; tail -3 $O/r2h_prefilter.err
timeout 600 python tools/bench_prefilter.py --rows 1000000 --normalize 1 >> $O/r2h_prefilter.jsonl 2>> $O/r2h_prefilter.err; echo "rc=$?"; tail -1 $O/r2h_prefilter.jsonl
timeout 600 python tools/bench_prefilter.py --rows 40000000 --d 384 --metric 1 >> $O/r2h_prefilter.jsonl 2>> $O/r2h_prefilter.err; echo "rc=$?"; tail -1 $O/r2h_prefilter.jsonl
