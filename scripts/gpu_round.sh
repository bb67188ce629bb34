#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3"
python bench.py > gpurun_out/bench_pcp_r3.json 2> gpurun_out/bench_pcp_r3.err; cut -c1-300 gpurun_out/bench_pcp_r3.json
python bench.py --rollout --no-cpu-baseline > gpurun_out/bench_pcp_rollout_r3.json 2> /dev/null
for s in Warehouse MaterialTransport ArcticTransport Simple; do python bench.py --scenario $s --envs 262144 --steps 100 --warmup 5 > gpurun_out/bench_${s}_r3.json 2> gpurun_out/bench_${s}_r3.err; cut -c1-200 gpurun_out/bench_${s}_r3.json; done
python bench.py --envs 131072 --steps 10 --warmup 3 $P20 > gpurun_out/bench_pcp20_r3.json 2> gpurun_out/bench_pcp20_r3.err; cut -c1-300 gpurun_out/bench_pcp20_r3.json
ncu --set full --import-source on --clock-control none -k regex:step_thread -s 3 -c 1 -o gpurun_out/ncu_pcp4_r3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pcp4_r3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_bench_pcp.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r3.json 2>/dev/null; cut -c1-300 gpurun_out/bench_ref_r3.json
