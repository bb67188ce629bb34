"""Generate tests/golden/episodes/*.npz: whole multi-episode trajectories of the UNMODIFIED reference
(run on oracle/shims) from a config seed, for the seed-for-seed reset tests.

ORACLE / TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference).
    python oracle/gen_episodes.py

Per fixture: the reference Wrapper is built with `seed` in its config (the scenario constructor seeds
numpy's global RNG with it), Python's `random` is seeded with py_seed (ArcticTransport draws its goal
column from it), then reset() / step(actions[t]) / reset() on done for T steps.  Stored per step: actions,
the reference's obs / reward / done / message / dist, whether a reset followed, the largest IPM iteration
count of the step's QP solves (steps with >= 25 are cvxopt limit cycles), and the state
right after every reset.
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_harness import RefEnv, _ensure_paths  # noqa: E402
from gen_golden import N_ACTIONS  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "episodes")


def _max_qp_iters(fn):
    _ensure_paths()
    import cvxopt.solvers as cs
    mx = [0]
    orig = cs.coneqp_l

    def wrapped(*a, **k):
        r = orig(*a, **k)
        mx[0] = max(mx[0], r["iterations"])
        return r
    cs.coneqp_l = wrapped
    try:
        out = fn()
    finally:
        cs.coneqp_l = orig
    return out, mx[0]


def episode_fixture(scenario, seed, steps, **over):
    py_seed = seed + 77
    env = RefEnv(scenario, seed=seed, **over)
    random.seed(py_seed)
    env.reset()
    resets = [env.get_state()]
    rng = np.random.RandomState(seed + 2000)          # actions: a private stream, not the global one
    acts, obs, rew, done, msg, dist, its, reset_after = [], [], [], [], [], [], [], []
    for _ in range(steps):
        a = rng.randint(0, N_ACTIONS[scenario], size=env.N)
        out, it = _max_qp_iters(lambda: env.step(a))
        acts.append(a), obs.append(out["obs"]), rew.append(out["reward"]), done.append(bool(out["done"][0]))
        msg.append(int(out["message"])), dist.append(out["dist"]), its.append(it)
        reset_after.append(bool(out["done"][0]))
        if out["done"][0]:
            env.reset()
            resets.append(env.get_state())
    blob = {"scenario": np.array(scenario), "cfg_json": np.array(json.dumps(env.cfg, sort_keys=True)),
            "py_seed": np.int64(py_seed), "actions": np.asarray(acts, dtype=np.int32),
            "obs": np.asarray(obs, dtype=np.float64), "reward": np.asarray(rew, dtype=np.float64),
            "done": np.asarray(done), "message": np.asarray(msg, dtype=np.int32),
            "dist": np.asarray(dist, dtype=np.float64), "qp_max_iters": np.asarray(its, dtype=np.int32),
            "reset_after": np.asarray(reset_after)}
    for k in resets[0]:
        blob["reset." + k] = np.stack([np.asarray(r[k]) for r in resets])
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "%s_seed%d.npz" % (scenario, seed))
    np.savez_compressed(path, **blob)
    print("%-40s %4d steps %3d episodes  stalls %d  %6.1f KB" % (
        os.path.basename(path), steps, len(resets), int(sum(i >= 25 for i in its)), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    for scn in ("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple"):
        for seed in (3, 11):
            episode_fixture(scn, seed, 150)
