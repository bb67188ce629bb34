#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 300 --warmup 5 > gpurun_out/bench_pcp_8gpu.json 2> gpurun_out/bench_pcp_8gpu.err; cut -c1-330 gpurun_out/bench_pcp_8gpu.json; tail -2 gpurun_out/bench_pcp_8gpu.err | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_ref_8gpu.json 2> gpurun_out/bench_ref_8gpu.err; cut -c1-200 gpurun_out/bench_ref_8gpu.json
