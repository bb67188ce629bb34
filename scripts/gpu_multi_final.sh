#!/bin/bash
# The multi-GPU bench lines of the final kernels (profiles/r02_bench_*_{2,4,8}gpu_final.json); scripts/gpu_multi.sh is the full
# 2 / 4 / 8 matrix of the middle of the round.  Usage: gpurun --gpus 8 -- 'bash scripts/gpu_multi_final.sh 8'   (or 4: config 5 at 4 and 2)
mkdir -p gpurun_out
run() {   # run <gpus> <tag> <bench args...>
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/multi_${tag}_${n}gpu.json 2> gpurun_out/multi_${tag}_${n}gpu.err
  tail -1 gpurun_out/multi_${tag}_${n}gpu.json | cut -c1-220
}
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3 --envs 131072"
if [ "${1:-8}" = 8 ]; then
  run 8 pcp20 $P20 --steps 8 --warmup 3
  run 8 pcp --steps 200
  run 8 wh --scenario Warehouse --envs 262144 --steps 50
  run 8 mt --scenario MaterialTransport --envs 262144 --steps 60
  run 8 at --scenario ArcticTransport --envs 262144 --steps 60
else
  run 4 pcp20 $P20 --steps 8 --warmup 3
  run 2 pcp20 $P20 --steps 8 --warmup 3
fi
