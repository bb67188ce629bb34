#!/bin/bash
mkdir -p gpurun_out
for d in -1 0 1; do MRB_HOST_DIRECT=$d python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_path or wrapper" 2>&1 | tail -1 | sed "s/^/direct=$d /"; done | tee gpurun_out/t_host.log
(for c in 2 4 8 16; do MRB_HOST_CHUNKS=$c python scripts/e2e_sweep.py $c 2>&1 | tail -1 | sed "s/^/hybrid /"; done
MRB_HOST_DIRECT=1 python scripts/e2e_sweep.py 4 2>&1 | tail -1 | sed "s/^/direct=1 /"
MRB_HOST_DIRECT=0 MRB_HOST_CHUNKS=4 python scripts/e2e_sweep.py 4 2>&1 | tail -1 | sed "s/^/direct=0 /") | tee gpurun_out/e2e_direct.log
python scripts/e2e_sweep.py 2>&1 | grep raw
