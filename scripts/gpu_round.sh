#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "Warehouse or MaterialTransport or capability" > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
for v in "" _v2 _v3 _v4 _v5; do
  echo "variant '$v'"
  MARBLER_B200_LIB=$PWD/marbler_b200/libmarbler_b200$v.so python scripts/quick_time.py Warehouse 262144 40 2>&1 | tail -1
done
