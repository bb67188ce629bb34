#!/bin/bash
mkdir -p gpurun_out
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3"
python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline $P20 > gpurun_out/bench_pcp20_nc.json 2>&1; python -c "
import json;d=json.load(open('gpurun_out/bench_pcp20_nc.json'));print('NC20', d['ms_per_step'])"
MRB_WARP_GENERIC=1 python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline $P20 > gpurun_out/bench_pcp20_gen.json 2>&1; python -c "
import json;d=json.load(open('gpurun_out/bench_pcp20_gen.json'));print('generic', d['ms_per_step'])"
(timeout 900 python -m pytest tests -m gpu -x -q -k "PCP20 or chunked or full_size") > gpurun_out/t.log 2>&1; grep -E "passed|failed" gpurun_out/t.log; grep -E "^E  |^FAILED" gpurun_out/t.log | head -10 | cut -c1-300
