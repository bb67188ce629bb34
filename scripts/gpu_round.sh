#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/lockstep_counts.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log; grep "accurate policy" gpurun_out/t_all.log
python scripts/policy_time.py 2>&1 | head -1
python scripts/policy_time.py accurate 2>&1 | head -1
python scripts/quick_time.py PredatorCapturePrey 65536 100 2>&1 | tail -1
python scripts/quick_time.py Warehouse 262144 50 2>&1 | tail -1
python bench.py > gpurun_out/bench_pcp.json 2> gpurun_out/bench_pcp.err; tail -1 gpurun_out/bench_pcp.json | cut -c1-250
python bench.py --scenario Warehouse --envs 262144 --steps 100 > gpurun_out/bench_wh.json 2> gpurun_out/bench_wh.err; tail -1 gpurun_out/bench_wh.json | cut -c1-200
