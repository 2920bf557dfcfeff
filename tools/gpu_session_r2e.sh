#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_shard_batch_gpu.py tests/test_parity_gpu.py -x -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
tail -30 $O/r2e_pytest.log
timeout 600 python tools/bench_streamed.py --rows 10000000 --d 768 --metric 0 --cap 65536 > $O/r2e_streamed_10M_forced.jsonl 2> $O/r2e_streamed_10M.err; echo "rc=$?"; cat $O/r2e_streamed_10M_forced.jsonl; tail -3 $O/r2e_streamed_10M.err
timeout 900 python tools/bench_streamed.py > $O/r2e_streamed_100M.jsonl 2> $O/r2e_streamed_100M.err; echo "rc=$?"; cat $O/r2e_streamed_100M.jsonl; tail -3 $O/r2e_streamed_100M.err
