#!/bin/bash
set -u
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/dur_pytest.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 )) s"
grep -A30 "slowest" gpurun_out/dur_pytest.log | head -34; tail -2 gpurun_out/dur_pytest.log
