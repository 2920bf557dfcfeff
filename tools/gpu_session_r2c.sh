#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_shard_batch_gpu.py tests/test_gemm_gpu.py -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
tail -30 $O/r2c_pytest.log
timeout 600 python bench.py --workload 10Mx768_ip_f32_k100_nq10000 --steps 5 --warmup 3 --no-cpu > $O/r2c_bench_cfg2.json 2> $O/r2c_bench_cfg2.err; echo "bench rc=$?"
tail -3 $O/r2c_bench_cfg2.err
