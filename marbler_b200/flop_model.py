"""FP64 flop model of the step kernels: useful floating-point operations of a launch from the counters the kernels
accumulate in the statistics vector (sub-steps executed, controller evaluations = QP solves, interior-point
iterations, env steps):

    flops = a * substeps + b * qp_solves + c * qp_iterations + d * env_steps

The coefficients are FITTED, per kernel, to instruction counts measured with ncu on the B200
(smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum and sm__inst_executed_pipe_tensor_subpipe_dmma.sum,
flops = dadd + dmul + 2 dfma + 512 dmma) over launches of
config variants that decorrelate the four counters (update_frequency 15 / 29 / 45, robotarium = True);
scripts/fp64_flop_model.py collects the data and does the fit, profiles/r02_fp64_flop_model.json holds the raw
counts, the fit and its residuals.  Thread-level predicated-on counts: lanes idling in a diverged warp do not
count, so these are the flops the envs needed, not the issue slots the warps spent.
bench.py multiplies the live counters of the timed region with these coefficients (roofline_fp64.achieved)."""

# (scenario, robots) -> (a, b, c, d); fit residuals <= 0.6 % over 24 launches per kernel (profiles/r02_fp64_flop_model.json)
COEFFICIENTS = {
    ("PredatorCapturePrey", 4): (96.3, 1656.8, 756.0, 220.5),
    ("Simple", 4): (96.3, 1656.9, 822.0, 120.6),
    ("MaterialTransport", 4): (95.8, 1292.9, 869.7, 114.5),
    ("ArcticTransport", 4): (94.5, 1108.8, 892.4, 153.8),
    ("Warehouse", 6): (176.0, 2309.2, 3114.6, 230.2),
    # one env per warp: the counts include the arithmetic every lane repeats (e.g. the 4 x 4 diagonal blocks of the
    # factorisation are computed by all 32 lanes) and 512 flops per DMMA.8x8x4 whether or not all eight columns of the
    # product are used (the tile solves use one): executed rather than minimal flops -- an occupancy figure of the FP64
    # units, not comparable with the count of the round-1 kernel (84.8 k per iteration, no DMMA)
    ("PredatorCapturePrey", 20): (1465.3, 103290.5, 143760.6, 8656.3),
}


def flops(scenario, n_robots, stats):
    co = COEFFICIENTS.get((scenario, int(n_robots)))
    if co is None:
        return None
    a, b, c, d = co
    total = a * stats["substeps"] + b * stats["qp_solves"] + c * stats["qp_iterations"] + d * stats["env_steps"]
    return {"total": total, "model": {"per_substep": a, "per_solve": b, "per_iteration": c, "per_env_step": d,
                                      "source": "profiles/r02_fp64_flop_model.json (fit to ncu instruction counts)"}}
