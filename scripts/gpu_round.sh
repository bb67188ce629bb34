#!/bin/bash
# scratch driver for one gpurun call
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/t.log 2>&1; grep -E "passed|failed" gpurun_out/t.log
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3"
python bench.py --envs 131072 --steps 10 --warmup 3 --no-cpu-baseline $P20 > gpurun_out/bench_pcp20.json 2> gpurun_out/bench_pcp20.err; cut -c1-330 gpurun_out/bench_pcp20.json
python bench.py --scenario Warehouse --envs 262144 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wh.json 2> gpurun_out/bench_wh.err; cut -c1-330 gpurun_out/bench_wh.json
