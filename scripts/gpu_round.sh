#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
for s in PredatorCapturePrey MaterialTransport ArcticTransport Simple Warehouse; do
  B=262144; [ $s = PredatorCapturePrey ] && B=65536
  python scripts/quick_time.py $s $B 100 2>&1 | tail -1
done
python scripts/quick_time.py PredatorCapturePrey 32768 5 predator=10 capture=10 ROBOT_INIT_RIGHT_THRESH=0.1 num_neighbors=3 2>&1 | tail -1
