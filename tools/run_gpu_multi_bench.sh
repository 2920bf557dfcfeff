#!/bin/bash
# N GPUs (gpurun --gpus N): the default bench line at N ranks (headline + other_configs), nothing else
set -u
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 100 --warmup 5 --no-cpu > $O/multi_bench_n$N.json 2> $O/multi_bench_n$N.err; echo "bench rc=$?"
tail -3 $O/multi_bench_n$N.err
