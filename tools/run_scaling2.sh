#!/bin/bash
mkdir -p gpurun_out
run() {
  if [ $1 -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu --steps $2 > gpurun_out/scale3_n$1.json 2> gpurun_out/scale3_n$1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29800+$1)) bench.py --gpus $1 --steps $2 > gpurun_out/scale3_n$1.json 2> gpurun_out/scale3_n$1.err
  fi
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/scale3_n$1.json").read().strip().splitlines()[-1])
    print("n$1", "QPS", round(j["value"],1), "ms", round(j["ms_per_step"],4), "scanGB/s", round(j["roofline"]["achieved"]), "e2e", round(j["e2e"]["value"],1), j["config"]["exchange"][:20])
except Exception as e:
    print("n$1 FAILED", e); print(open("gpurun_out/scale3_n$1.err").read()[-600:])
PY
}
run 1 200
run 2 300
run 4 300
run 8 400
timeout 300 python -m pytest tests/test_sharded_nccl_gpu.py -q -m gpu --timeout 250 2>&1 | tail -1
