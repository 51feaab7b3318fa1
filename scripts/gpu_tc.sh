#!/bin/bash
# Tensor-core kernel bring-up: unit tests first (tight timeout so a deadlocked mbarrier cannot eat the budget).
set -u
mkdir -p gpurun_out
timeout -k 10 ${TC_TIMEOUT:-300} python -m pytest tests/test_gpu_conv_tc.py -q --timeout 120 ${TC_ARGS:-} 2>&1 | grep -vE "^(shape|scale|shift|kernel|x0|x1|res|ref|got) =|^\s+\[|dtype=float32" > gpurun_out/pytest_tc.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_tc.log
grep -E "passed|failed|Error|error|exit|max abs err" gpurun_out/pytest_tc.log | tail -40
