#!/bin/bash
# scratch runner for one gpurun call
mkdir -p gpurun_out
rm -f gpurun_out/lockstep_counts.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for s in PredatorCapturePrey Warehouse MaterialTransport ArcticTransport Simple; do
  B=262144; [ $s = PredatorCapturePrey ] && B=65536
  python scripts/quick_time.py $s $B 50 2>&1 | tail -1
done
python scripts/quick_time.py PredatorCapturePrey 32768 5 predator=10 capture=10 ROBOT_INIT_RIGHT_THRESH=0.1 num_neighbors=3 2>&1 | tail -1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency scripts/microbench/fp64_latency.cu && /tmp/fp64_latency > gpurun_out/fp64_latency.txt 2>&1; head -8 gpurun_out/fp64_latency.txt
python bench.py > gpurun_out/bench_pcp.json 2> gpurun_out/bench_pcp.err; tail -1 gpurun_out/bench_pcp.json | cut -c1-300
