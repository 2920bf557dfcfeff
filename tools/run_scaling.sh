#!/bin/bash
# 1/2/4/8-GPU strong scaling of the headline workload (run under gpurun --gpus 8)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu --steps 200 > gpurun_out/scale_r1_n$n.json 2> gpurun_out/scale_r1_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 200 > gpurun_out/scale_r1_n$n.json 2> gpurun_out/scale_r1_n$n.err
  fi
  tail -c 400 gpurun_out/scale_r1_n$n.json | head -c 400; echo
done
for n in 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $n --steps 100 --workload 100Mx384_l2_f32_k10_nq1 > gpurun_out/scale_r1_100M_n$n.json 2> gpurun_out/scale_r1_100M_n$n.err
  tail -c 300 gpurun_out/scale_r1_100M_n$n.json; echo
done
