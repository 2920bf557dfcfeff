#!/bin/bash
# 1/2/4/8-GPU strong scaling of the headline workload (run under gpurun --gpus 8), fused vs NCCL exchange
mkdir -p gpurun_out
run() { # n exchange workload steps tag
  if [ $1 -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu --steps $4 --workload $3 > gpurun_out/scale_r1_$5.json 2> gpurun_out/scale_r1_$5.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500+$1)) bench.py --gpus $1 --steps $4 --exchange $2 --workload $3 > gpurun_out/scale_r1_$5.json 2> gpurun_out/scale_r1_$5.err
  fi
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/scale_r1_$5.json").read().strip().splitlines()[-1])
    print("$5", "QPS", round(j["value"],1), "ms", round(j["ms_per_step"],4), "scanGB/s", round(j["roofline"]["achieved"]), "e2e", round(j["e2e"]["value"],1), "launches", j["gpu_launches"])
except Exception as e:
    print("$5 FAILED", e); print(open("gpurun_out/scale_r1_$5.err").read()[-800:])
PY
}
W=10Mx768_ip_f32_k10_nq1
run 1 fused $W 200 n1
run 2 fused $W 200 n2_fused
run 4 fused $W 200 n4_fused
run 8 fused $W 300 n8_fused
run 8 nccl $W 300 n8_nccl
run 4 nccl $W 200 n4_nccl
run 8 fused 100Mx384_l2_f32_k10_nq1 100 100M_n8_fused
