#!/usr/bin/env python
"""Fit the FP64 flop model of the step kernels (marbler_b200/flop_model.py) to ncu instruction counts.

    # on the B200 box, ONE GPU (ncu serialises and replays every kernel):
    ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,gpu__time_duration.sum \
        --clock-control none -k regex:step_ \
        --csv --log-file gpurun_out/fp64_counts.csv python scripts/fp64_flop_model.py collect
    # anywhere:
    python scripts/fp64_flop_model.py fit gpurun_out/fp64_counts.csv gpurun_out/fp64_launches.json profiles/r02_fp64_flop_model.json

`collect` steps every workload through config variants that decorrelate the four counters of the model
(sub-steps, QP solves, IPM iterations, env steps): update_frequency 15 / 29 / 45 and robotarium = True, and records
the statistics-vector delta of every launch (same order as ncu's launch list).  `fit` joins the two, solves the
least-squares problem per kernel and prints the coefficients to paste into flop_model.COEFFICIENTS.
flops = dadd + dmul + 2 dfma, thread-level, predicated-on (lanes idling in a diverged warp do not count), + 512 per warp-level
DMMA.8x8x4 (the 20-robot kernel's tile updates and tile solves; 8 x 8 x 4 multiply-adds, executed whether or not every column
of the product is used)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PCP20 = dict(predator=10, capture=10, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3)
WORKLOADS = [("PredatorCapturePrey", {}, 8192), ("Warehouse", {}, 8192), ("MaterialTransport", {}, 8192),
             ("ArcticTransport", {}, 8192), ("Simple", {}, 8192), ("PredatorCapturePrey", PCP20, 1024)]
VARIANTS = [{}, {"update_frequency": 15}, {"update_frequency": 45}, {"robotarium": True, "update_frequency": 10}]
STEPS = 6


def collect(out_path):
    import torch
    from marbler_b200 import config
    from marbler_b200.vec_env import VecEnv
    records = []
    for scenario, over, B in WORKLOADS:
        for var in VARIANTS:
            cfg = config.load_yaml(config.default_config_path(scenario))
            cfg.update(over)
            cfg.update(var)
            env = VecEnv(scenario, cfg, num_envs=B, device="cuda:0", seed=3, auto_reset=True)
            env.reset()
            gen = torch.Generator(device="cuda:0").manual_seed(5)
            prev = env.read_stats()
            for t in range(STEPS):
                a = torch.randint(0, env.n_actions, (B, env.N), generator=gen, device="cuda:0", dtype=torch.int32)
                env.step(a)
                torch.cuda.synchronize()
                now = env.read_stats()
                records.append({"scenario": scenario, "robots": env.N, "variant": var, "step": t, "envs": B,
                                "delta": {k: now[k] - prev[k] for k in now}})
                prev = now
            del env
    with open(out_path, "w") as f:
        json.dump(records, f)
    print("wrote %d launch records to %s" % (len(records), out_path))


def fit(csv_path, launches_path, out_path):
    import numpy as np
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    head = rows[0]
    ix = {k: head.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
    per = {}
    for r in rows[1:]:
        if "step_" not in r[ix["Kernel Name"]]:
            continue
        per.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    launches = [per[k] for k in sorted(per)]
    records = json.load(open(launches_path))
    assert len(launches) == len(records), (len(launches), len(records))
    groups = {}
    for l, r in zip(launches, records):
        fl = l["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"] + l["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] \
            + 2.0 * l["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + 512.0 * l.get("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", 0.0)
        d = r["delta"]
        groups.setdefault((r["scenario"], r["robots"]), []).append(
            ([d["substeps"], d["qp_solves"], d["qp_iterations"], d["env_steps"]], fl, l, r))
    out = {"how": __doc__, "kernels": {}}
    for (scenario, robots), items in sorted(groups.items()):
        A = np.array([i[0] for i in items], dtype=np.float64)
        y = np.array([i[1] for i in items], dtype=np.float64)
        co, *_ = np.linalg.lstsq(A, y, rcond=None)
        rel = np.abs(A @ co - y) / y
        print("(%r, %d): (%.1f, %.1f, %.1f, %.1f),   # max residual %.2f %%, %d launches" % (
            scenario, robots, co[0], co[1], co[2], co[3], 100 * rel.max(), len(items)))
        out["kernels"]["%s/%d" % (scenario, robots)] = {
            "coefficients": {"per_substep": co[0], "per_solve": co[1], "per_iteration": co[2], "per_env_step": co[3]},
            "max_relative_residual": float(rel.max()),
            "launches": [{"kernel": i[2]["kernel"], "variant": i[3]["variant"], "step": i[3]["step"], "envs": i[3]["envs"],
                          "counters": dict(zip(("substeps", "qp_solves", "qp_iterations", "env_steps"), i[0])),
                          "dadd": i[2]["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"],
                          "dmul": i[2]["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"],
                          "dfma": i[2]["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"],
                          "dmma_warp": i[2].get("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", 0.0),
                          "flops": i[1], "model_flops": float(np.dot(i[0], co)),
                          "ncu_duration_ns": i[2].get("gpu__time_duration.sum")} for i in items]}
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "collect":
        collect(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "fp64_launches.json"))
    else:
        fit(*sys.argv[2:5])
