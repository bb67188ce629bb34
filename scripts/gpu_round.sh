#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_policy.py -m gpu -x -q -k "matches_reference or fresh_mask or device_rollout" > gpurun_out/t_policy.log 2>&1; tail -2 gpurun_out/t_policy.log | cut -c1-300
python scripts/policy_time.py 2>&1 | grep "policy kernel\|rror" | head -3
