#!/bin/bash
mkdir -p gpurun_out
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3"
ncu --set full --import-source on --clock-control none -k regex:step_warp -c 1 -o gpurun_out/ncu_pcp20c python bench.py --envs 16384 --steps 1 --warmup 3 --no-cpu-baseline $P20 > gpurun_out/ncu_pcp20c.log 2>&1
