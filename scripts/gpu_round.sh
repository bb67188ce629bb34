#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_policy.py tests/test_epymarl_adapter.py -m gpu -x -q > gpurun_out/t_policy.log 2>&1; tail -3 gpurun_out/t_policy.log | cut -c1-300
python scripts/policy_time.py 2>&1 | grep "policy kernel\|rror" | head -3
MRB_POLICY_TC=0 python scripts/policy_time.py 2>&1 | grep "policy kernel\|rror" | head -3
/usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/san_memcheck.log
/usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/san_racecheck.log
ncu --set full --import-source on --clock-control none -k regex:policy_act_tc2 -s 3 -c 1 -o gpurun_out/ncu_policy_tc2c python scripts/policy_time.py > gpurun_out/ncu_policy_tc2c.log 2>&1
