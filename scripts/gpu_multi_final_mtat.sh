#!/bin/bash
# one gpurun --gpus 8 call with the final kernels: config 4 (MaterialTransport, ArcticTransport; 262,144 envs per GPU)
mkdir -p gpurun_out
run() {   # run <gpus> <tag> <bench args...>
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/multi_${tag}_${n}gpu.json 2> gpurun_out/multi_${tag}_${n}gpu.err
  tail -1 gpurun_out/multi_${tag}_${n}gpu.json | cut -c1-220
}
run 8 mt --scenario MaterialTransport --envs 262144 --steps 60
run 8 at --scenario ArcticTransport --envs 262144 --steps 60
