#!/bin/bash
# one gpurun --gpus 4 call with the final kernels: config 5 (20 robots, 131,072 envs per GPU) at 4 and 2 GPUs
mkdir -p gpurun_out
run() {   # run <gpus> <tag> <bench args...>
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/multi_${tag}_${n}gpu.json 2> gpurun_out/multi_${tag}_${n}gpu.err
  tail -1 gpurun_out/multi_${tag}_${n}gpu.json | cut -c1-220
}
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3 --envs 131072"
run 4 pcp20 $P20 --steps 8 --warmup 3
run 2 pcp20 $P20 --steps 8 --warmup 3
