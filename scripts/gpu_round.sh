#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or rollout_lockstep or chunked or sharded or full_size_properties" > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
for s in PredatorCapturePrey MaterialTransport ArcticTransport Simple; do
  B=262144; [ $s = PredatorCapturePrey ] && B=65536
  python scripts/quick_time.py $s $B 50 2>&1 | tail -1
  MRB_SORT_ENVS=0 python scripts/quick_time.py $s $B 50 2>&1 | tail -1
done
python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench_pcp.json 2> gpurun_out/bench_pcp.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pcp.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["fp64"], d["roofline_fp64"]["frac"], d["roofline_fp64"]["peak"])
PY
MRB_SORT_ENVS=0 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench_pcp_nosort.json 2> gpurun_out/bench_pcp.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pcp_nosort.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["fp64"], d["roofline_fp64"]["frac"], d["roofline_fp64"]["peak"])
PY
