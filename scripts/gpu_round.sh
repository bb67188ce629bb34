#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_qp_crosscheck.py -m gpu -x -q -k "barrier_qp or team_sizes or fixture or full_size or crosscheck or qp" > gpurun_out/t_w20.log 2>&1; tail -2 gpurun_out/t_w20.log
P20="predator=10 capture=10 ROBOT_INIT_RIGHT_THRESH=0.1 num_neighbors=3"
python scripts/quick_time.py PredatorCapturePrey 32768 5 $P20 2>&1 | tail -1
python scripts/quick_time.py PredatorCapturePrey 32768 5 predator=4 capture=5 ROBOT_INIT_RIGHT_THRESH=0.1 num_neighbors=3 2>&1 | tail -1
