#!/bin/bash
# one gpurun --gpus 8 call: the multi-GPU bench lines of BASELINE.json configs 2, 4 and 5 (weak scaling, envs per GPU fixed)
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/topo.txt
run() {   # run <gpus> <tag> <bench args...>
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/multi_${tag}_${n}gpu.json 2> gpurun_out/multi_${tag}_${n}gpu.err
  python - "$tag" "$n" <<'PY'
import json, sys
tag, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open('gpurun_out/multi_%s_%sgpu.json' % (tag, n)).read().strip().splitlines()[-1])
    print(tag, n, "ms/step %.4f value %.4g e2e %.4g floor %.4g gbs %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"],
          d["e2e"]["pcie_floor"]["env_steps_per_s_at_floor"], d["e2e"]["pcie_floor"]["aggregate_gbs"]), d["e2e"].get("cpu_affinity"))
except Exception as ex:
    print(tag, n, "FAILED", ex)
PY
}
P20="--override predator=10 --override capture=10 --override ROBOT_INIT_RIGHT_THRESH=0.1 --override num_neighbors=3 --envs 131072"
for n in 8 4 2; do run $n pcp --steps 200; done
run 8 pcp_nobind --steps 200 --no-numa-bind
for n in 8 4 2; do run $n mt --scenario MaterialTransport --envs 262144 --steps 100; done
for n in 8 4 2; do run $n at --scenario ArcticTransport --envs 262144 --steps 100; done
for n in 8 4 2; do run $n pcp20 $P20 --steps 8 --warmup 3; done
run 8 wh --scenario Warehouse --envs 262144 --steps 50
