#!/bin/bash
mkdir -p gpurun_out
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3"
python bench.py --envs 131072 --steps 10 --warmup 3 $P20 > gpurun_out/bench_pcp20_r5.json 2> gpurun_out/bench_pcp20_r5.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_pcp20_r5.json').read().strip().split('\n')[-1]); print('pcp20', d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
ncu --set full --import-source on --clock-control none -k regex:step_warp -c 1 -o gpurun_out/ncu_pcp20e python bench.py --envs 16384 --steps 1 --warmup 3 --no-cpu-baseline $P20 > gpurun_out/ncu_pcp20e.log 2>&1
