#!/bin/bash
# N GPUs (gpurun --gpus N): multi-GPU parity tests, config-2 (batched) bench and the headline bench at N ranks
set -u
O=gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_sharded_nccl_gpu.py -x -q > $O/multi_pytest_n$N.log 2>&1; echo "pytest rc=$?" >> $O/multi_pytest_n$N.log
tail -15 $O/multi_pytest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --workload 10Mx768_ip_f32_k100_nq10000 --steps 5 --warmup 3 --no-cpu > $O/multi_bench_cfg2_n$N.json 2> $O/multi_bench_cfg2_n$N.err; echo "cfg2 rc=$?"
tail -3 $O/multi_bench_cfg2_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-cpu > $O/multi_bench_n$N.json 2> $O/multi_bench_n$N.err; echo "bench rc=$?"
tail -3 $O/multi_bench_n$N.err
