#!/bin/bash
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $S --tool memcheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "^ok" gpurun_out/san_memcheck.log; tail -3 gpurun_out/san_memcheck.log
timeout 900 $S --tool racecheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "^ok" gpurun_out/san_racecheck.log; tail -3 gpurun_out/san_racecheck.log
MRB_POLICY_TC=1 timeout 600 $S --tool memcheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/san_memcheck_tc.log 2>&1; echo "memcheck tc rc=$?"; tail -3 gpurun_out/san_memcheck_tc.log
