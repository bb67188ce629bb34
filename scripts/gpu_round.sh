#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t.log 2>&1; grep -E "passed|failed" gpurun_out/t.log; grep -E "^E  |^FAILED" gpurun_out/t.log | head -20
python bench.py --scenario Warehouse --envs 262144 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wh.json 2> gpurun_out/bench_wh.err; cut -c1-200 gpurun_out/bench_wh.json
