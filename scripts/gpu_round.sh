#!/bin/bash
mkdir -p gpurun_out
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3"
python bench.py > gpurun_out/bench_pcp_r4.json 2> gpurun_out/bench_pcp_r4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_pcp_r4.json').read().strip().split('\n')[-1]); print('pcp', d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])"
for s in Warehouse MaterialTransport ArcticTransport Simple; do python bench.py --scenario $s --envs 262144 --steps 100 --warmup 5 > gpurun_out/bench_${s}_r4.json 2> gpurun_out/bench_${s}_r4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${s}_r4.json').read().strip().split('\n')[-1]); print('$s', d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'])"; done
python bench.py --envs 131072 --steps 10 --warmup 3 $P20 > gpurun_out/bench_pcp20_r4.json 2> gpurun_out/bench_pcp20_r4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_pcp20_r4.json').read().strip().split('\n')[-1]); print('pcp20', d['ms_per_step'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
