#!/bin/bash
# final measurement run of the round: tests, flop-model counts, ncu captures of the final kernels, bench lines of every config
mkdir -p gpurun_out
rm -f gpurun_out/lockstep_counts.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:step_ --csv --log-file gpurun_out/fp64_counts.csv python scripts/fp64_flop_model.py collect > gpurun_out/fp64_collect.log 2>&1; tail -1 gpurun_out/fp64_collect.log
ncu --set full --clock-control none --import-source on -k regex:step_thread -s 8 -c 1 -o gpurun_out/r02_pcp4_final -f python scripts/quick_time.py PredatorCapturePrey 65536 5 > gpurun_out/ncu_pcp4.log 2>&1; tail -1 gpurun_out/ncu_pcp4.log
ncu --set full --clock-control none --import-source on -k regex:step_thread -s 8 -c 1 -o gpurun_out/r02_wh6_final -f python scripts/quick_time.py Warehouse 262144 5 > gpurun_out/ncu_wh6.log 2>&1; tail -1 gpurun_out/ncu_wh6.log
ncu --set full --clock-control none --import-source on -k regex:step_warp -s 3 -c 1 -o gpurun_out/r02_pcp20_final -f python scripts/quick_time.py PredatorCapturePrey 16384 2 predator=10 capture=10 ROBOT_INIT_RIGHT_THRESH=0.1 num_neighbors=3 > gpurun_out/ncu_pcp20.log 2>&1; tail -1 gpurun_out/ncu_pcp20.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_pcp.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python bench.py > gpurun_out/bench_pcp.json 2> gpurun_out/bench_pcp.err; tail -1 gpurun_out/bench_pcp.json | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json; cut -c1-160 gpurun_out/bench_reference.json
for s in Warehouse MaterialTransport ArcticTransport Simple; do
  python bench.py --scenario $s --envs 262144 --steps 100 --cpu-seconds 5 > gpurun_out/bench_$s.json 2> gpurun_out/bench_$s.err; tail -1 gpurun_out/bench_$s.json | cut -c1-160
done
python bench.py --override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3 --envs 131072 --steps 10 --cpu-seconds 5 > gpurun_out/bench_pcp20.json 2> gpurun_out/bench_pcp20.err; tail -1 gpurun_out/bench_pcp20.json | cut -c1-160
python bench.py --rollout --steps 320 --no-cpu-baseline > gpurun_out/bench_rollout.json 2> gpurun_out/bench_rollout.err; tail -1 gpurun_out/bench_rollout.json | cut -c1-100
