#!/bin/bash
# scratch runner for one gpurun call
mkdir -p gpurun_out
rm -f gpurun_out/lockstep_counts.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
for s in PredatorCapturePrey Warehouse MaterialTransport ArcticTransport Simple; do
  B=262144; [ $s = PredatorCapturePrey ] && B=65536
  python scripts/quick_time.py $s $B 50 2>&1 | tail -1
done
