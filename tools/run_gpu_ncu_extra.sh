#!/bin/bash
# extra ncu captures: the scan kernel over the bf16 shadow (option prefilter) and the text embedder (K6)
set -u
O=gpurun_out
P="python tools/bench_prefilter.py --steps 6"
timeout 300 $P > $O/ncux_plain1.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 20 -c 2 -f -o $O/ncux_prefilter_scan $P > $O/ncux_1.log 2>&1
tail -2 $O/ncux_1.log
R="python tools/bench_rebuild.py --n 2000000 --py-n 100000 --host-n 50000"
timeout 300 $R > $O/ncux_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:text_embed -s 2 -c 2 -f -o $O/ncux_text_embed $R > $O/ncux_2.log 2>&1
tail -2 $O/ncux_2.log
