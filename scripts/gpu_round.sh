#!/bin/bash
mkdir -p gpurun_out
for v in "" _whL _whM _whO; do
  echo "variant '$v'"
  MARBLER_B200_LIB=$PWD/marbler_b200/libmarbler_b200$v.so python scripts/quick_time.py Warehouse 262144 30 2>&1 | tail -1
done 2>&1 | tee gpurun_out/wh_variants3.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "barrier_qp or Warehouse or team_sizes" > gpurun_out/t_wh.log 2>&1; tail -3 gpurun_out/t_wh.log
