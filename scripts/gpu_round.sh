#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_policy.py -m gpu -q) > gpurun_out/t.log 2>&1; grep -E "passed|failed" gpurun_out/t.log; grep -E "^E  |^FAILED|Error" gpurun_out/t.log | head -30 | cut -c1-300
timeout 120 python scripts/policy_time.py 2>&1 | grep policy
