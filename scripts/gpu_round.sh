#!/bin/bash
# scratch runner for one gpurun call: full GPU test suite, the default bench line and the reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_pcp.json 2> gpurun_out/bench_pcp.err; tail -1 gpurun_out/bench_pcp.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
