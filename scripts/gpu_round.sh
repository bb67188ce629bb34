#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python bench.py --rollout --no-cpu-baseline --steps 320 > gpurun_out/bench_pcp_rollout_r5.json 2> /dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_pcp_rollout_r5.json').read().strip().split('\n')[-1]); print(d['rollout'])"
