#!/bin/bash
# run the test files given as arguments on the GPU box
set -u
timeout 900 python -m pytest "$@" -x -q > gpurun_out/one_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/one_pytest.log
