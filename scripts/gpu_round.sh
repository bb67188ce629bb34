#!/bin/bash
mkdir -p gpurun_out
for v in _W5 _W6 _W8; do
  MARBLER_B200_LIB=$PWD/marbler_b200/libmarbler_b200$v.so python bench.py --scenario Warehouse --envs 262144 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wh$v.json 2>/dev/null; echo "V$v"; python -c "
import json;d=json.load(open('gpurun_out/bench_wh$v.json'));print(d['ms_per_step'])"
done
