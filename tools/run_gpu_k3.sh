#!/bin/bash
# one GPU box visit for the batched path: config-2 timing with clocks + cuBLAS beside it, small batches, the K3 test
# files, and an ncu --set full capture of gemm_topk_kernel (after its plain run exited 0)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/bench_batched.py --reps 3 > $O/k3_batched.jsonl 2> $O/k3_batched.err; echo "bench_batched rc=$?"
timeout 120 python tools/bench_small_batches.py > $O/k3_small_batches.jsonl 2>&1; echo "small_batches rc=$?"
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_shard_batch_gpu.py tests/test_exact_integer_gpu.py tests/test_prefilter_gpu.py -x -q -m gpu > $O/k3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/k3_pytest.log
K="python tools/k3_probe.py 2000000 768 4096 100"
timeout 300 $K > $O/k3_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 2 -f -o $O/k3_prof_gemm $K > $O/k3_ncu.log 2>&1
tail -2 $O/k3_ncu.log
