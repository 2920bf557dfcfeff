#!/bin/bash
set -u
O=gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_pytest.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 )) s" >> $O/tests_pytest.log
tail -6 $O/tests_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/tests_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/tests_smoke.log
