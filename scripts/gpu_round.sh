#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 > gpurun_out/bench_pcp_2gpu_r4.json 2> gpurun_out/bench_pcp_2gpu_r4.err; tail -1 gpurun_out/bench_pcp_2gpu_r4.json | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
