#!/bin/bash
# scratch runner for one gpurun call: full GPU test suite, smoke, the default bench line, the reference arm,
# per-config bench lines and the ncu instruction counts the FP64 flop model is fitted to
mkdir -p gpurun_out
rm -f gpurun_out/lockstep_counts.jsonl
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_pcp.json 2> gpurun_out/bench_pcp.err; tail -1 gpurun_out/bench_pcp.json | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:step_ --csv --log-file gpurun_out/fp64_counts.csv python scripts/fp64_flop_model.py collect > gpurun_out/fp64_collect.log 2>&1; tail -1 gpurun_out/fp64_collect.log
python bench.py --scenario Warehouse --envs 262144 --steps 100 --no-cpu-baseline > gpurun_out/bench_wh.json 2> gpurun_out/bench_wh.err; tail -1 gpurun_out/bench_wh.json | cut -c1-200
python bench.py --override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3 --envs 131072 --steps 10 --no-cpu-baseline > gpurun_out/bench_pcp20.json 2> gpurun_out/bench_pcp20.err; tail -1 gpurun_out/bench_pcp20.json | cut -c1-200
