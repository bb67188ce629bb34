"""Generate tests/golden/*.npz by running /root/reference/robotarium_gym UNMODIFIED on oracle/shims.

ORACLE / TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference).
    python oracle/gen_golden.py            # rewrites every fixture (deterministic: fixed seeds)

Every fixture is a set of independent one-step cases: pre-state s0, actions, the reference's outputs
and post-state s1 (RefEnv.step_from).  "rollout" sets take s0 from the reference's own trajectory
(reset distribution + random actions, resets on done); "inject" sets use synthetic states that force
the rare events (collisions, boundary exits, captures, loading, goal cells, time-outs).
The reference's QP is the restated cvxopt in oracle/shims (PARITY UNPINNED vs real cvxopt/rps).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_harness import RefEnv, last_qp_iterations, _ensure_paths  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
N_ACTIONS = {"PredatorCapturePrey": 5, "Warehouse": 5, "MaterialTransport": 20, "ArcticTransport": 5, "Simple": 5}

PCP20 = dict(predator=10, capture=10, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3, n_agents=20)


def _stack(dicts):
    return {k: np.stack([np.asarray(d[k]) for d in dicts]) for k in dicts[0]}


def save(name, env, s0, actions, outs, s1, qp_iters):
    os.makedirs(OUT, exist_ok=True)
    blob = {"scenario": np.array(env.scenario), "cfg_json": np.array(json.dumps(env.cfg, sort_keys=True)),
            "actions": np.asarray(actions, dtype=np.int32), "qp_iters": np.asarray(qp_iters, dtype=np.int32)}
    for pre, ds in (("s0.", s0), ("s1.", s1), ("out.", outs)):
        for k, v in _stack(ds).items():
            blob[pre + k] = v
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print("%-34s %4d cases  viol %3d  done %3d  %7.1f KB" % (
        name, len(actions), int(sum(int(o["message"]) != 0 for o in outs)),
        int(sum(bool(o["done"][0]) for o in outs)), os.path.getsize(path) / 1024.0))


def _count_qp_iters(fn):
    """Total IPM iterations of all QPs solved inside fn() (shim statistic)."""
    _ensure_paths()
    import cvxopt.solvers as cs
    total = [0]
    orig = cs.coneqp_l

    def wrapped(*a, **k):
        r = orig(*a, **k)
        total[0] += r["iterations"]
        return r
    cs.coneqp_l = wrapped
    try:
        out = fn()
    finally:
        cs.coneqp_l = orig
    return out, total[0]


def rollout(name, scenario, steps, seed, **over):
    np.random.seed(seed)
    import random
    random.seed(seed)
    env = RefEnv(scenario, seed=seed, **over)
    env.reset()
    rng = np.random.RandomState(seed + 1000)
    s0, acts, outs, s1, its = [], [], [], [], []
    for _ in range(steps):
        st = env.get_state()
        a = rng.randint(0, N_ACTIONS[scenario], size=env.N)
        out, it = _count_qp_iters(lambda: env.step(a))
        s0.append(st), acts.append(a), outs.append(out), s1.append(env.get_state()), its.append(it)
        if out["done"][0]:
            env.reset()
    save(name, env, s0, acts, outs, s1, its)


# ------------------------------------------------------------------ synthetic states
def _poses(rng, N, crowd):
    """Random poses; `crowd` = probability of forcing a close pair / an out-of-bounds robot."""
    p = np.empty((3, N))
    for _ in range(200):                      # large teams: start from a collision-free layout
        p[0] = rng.uniform(-1.55, 1.55, N)
        p[1] = rng.uniform(-0.95, 0.95, N)
        d = np.hypot(p[0][:, None] - p[0][None], p[1][:, None] - p[1][None]) + 9 * np.eye(N)
        if N < 10 or d.min() > 0.16:
            break
    p[2] = rng.uniform(-np.pi, np.pi, N)
    if rng.rand() < crowd:                    # a pair near the collision / barrier radii
        i, j = rng.choice(N, 2, replace=False)
        ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(0.10, 0.30)
        p[0, j] = np.clip(p[0, i] + d * np.cos(ang), -1.58, 1.58)
        p[1, j] = np.clip(p[1, i] + d * np.sin(ang), -0.98, 0.98)
    if rng.rand() < crowd * 0.5:              # somebody hugging / crossing the arena boundary
        i = rng.randint(N)
        if rng.rand() < 0.5:
            p[0, i] = rng.choice([-1, 1]) * rng.uniform(1.56, 1.63)
        else:
            p[1, i] = rng.choice([-1, 1]) * rng.uniform(0.96, 1.03)
    return p


def synth_state(env, rng, crowd=0.35):
    s, N, cfg = env.scenario, env.N, env.cfg
    st = {"poses": _poses(rng, N, crowd)}
    st["prev_valid"] = np.int32(rng.rand() < 0.8)
    st["prev_pose"] = st["poses"] + rng.normal(0, 0.004, (3, N)) if st["prev_valid"] else np.zeros((3, N))
    mx = cfg["max_episode_steps"]
    st["episode_steps"] = np.int32(rng.choice([0, 1, mx - 1, mx, rng.randint(0, mx + 1)]))
    if s == "PredatorCapturePrey":
        P = cfg["num_prey"]
        prey = np.stack([rng.choice([.5, .7, .9, 1.1], P), rng.uniform(-.9, .9, P)], 1)
        for p in range(P):                    # park robots on top of / near some prey
            if rng.rand() < 0.5:
                i = rng.randint(N)
                ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(0.0, 0.5)
                prey[p] = np.clip(st["poses"][:2, i] + d * np.array([np.cos(ang), np.sin(ang)]), -1.4, 1.4)
        sensed = rng.rand(P) < 0.4
        captured = sensed & (rng.rand(P) < 0.4)
        if rng.rand() < 0.1:
            captured[:] = True
            captured[rng.randint(P)] = False
            sensed = sensed | captured
        st.update(prey_loc=prey, prey_sensed=sensed.astype(np.uint8), prey_captured=captured.astype(np.uint8))
    elif s == "Warehouse":
        st["loaded"] = (rng.rand(N) < 0.5).astype(np.uint8)
        for i in range(N):                    # many robots inside the load / unload strips
            if rng.rand() < 0.5:
                st["poses"][0, i] = rng.choice([-1, 1]) * rng.uniform(0.9, 1.45)
    elif s == "MaterialTransport":
        st["load"] = rng.choice([0, 0, 3, 5, 15], N).astype(np.int32)
        st["zone_load"] = np.array([rng.choice([0, 4, 12, 100]), rng.choice([0, 3, 7, 20])], dtype=np.int32)
        st["messages"] = rng.randint(0, 4, 4).astype(np.int32)
        for i in range(N):
            r = rng.rand()
            if r < 0.3:
                st["poses"][0, i] = rng.choice([-1, 1]) * rng.uniform(0.9, 1.45)
            elif r < 0.5:
                ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(0.0, 0.45)
                st["poses"][:2, i] = d * np.cos(ang), d * np.sin(ang)
    elif s == "ArcticTransport":
        grid = rng.randint(0, 3, (8, 12))
        g = rng.randint(1, 12)
        grid[0:2, g - 1:g + 1] = 3
        grid[7, 1:11] = 0
        st.update(grid=grid.astype(np.uint8), goal_col=np.int32(g),
                  pixel_type=rng.randint(0, 4, N).astype(np.int32),
                  reached_goal=(rng.rand(N) < 0.3).astype(np.uint8))
        for i in range(N):                    # near the goal block now and then
            if rng.rand() < 0.3:
                st["poses"][0, i] = np.clip(g * .25 - 1.5 + rng.uniform(-.3, .3), -1.55, 1.55)
                st["poses"][1, i] = rng.uniform(0.4, 0.97)
    elif s == "Simple":
        st["goal"] = np.array([rng.choice([.5, .7, .9, 1.1]), rng.uniform(-.9, .9)])
    return st


def inject(name, scenario, cases, seed, crowd=0.35, **over):
    np.random.seed(seed)
    env = RefEnv(scenario, seed=seed, **over)
    env.reset()
    rng = np.random.RandomState(seed + 2000)
    s0, acts, outs, s1, its = [], [], [], [], []
    for _ in range(cases):
        st = synth_state(env, rng, crowd)
        a = rng.randint(0, N_ACTIONS[scenario], size=env.N)
        if scenario == "PredatorCapturePrey" and rng.rand() < 0.5:
            a[rng.rand(env.N) < 0.5] = 4      # 'no_action' is the capture action
        (out, post), it = _count_qp_iters(lambda: env.step_from(st, a))
        # store the state exactly as the reference saw it (set_state normalises dtypes)
        s0.append(st), acts.append(a), outs.append(out), s1.append(post), its.append(it)
    save(name, env, s0, acts, outs, s1, its)


def qp_vectors():
    """Barrier-certificate QP alone: (dxi, xi) -> u through rps' certificate on the restated cvxopt."""
    _ensure_paths()
    from rps.utilities.barrier_certificates import (create_single_integrator_barrier_certificate,
                                                    create_single_integrator_barrier_certificate2)
    safe = create_single_integrator_barrier_certificate2(safety_radius=.2)      # utilities/controller.py:14
    default = create_single_integrator_barrier_certificate()                    # utilities/controller.py:16
    rng = np.random.RandomState(7)
    blob = {}
    for N, cases in ((2, 64), (3, 64), (4, 512), (6, 256), (10, 64), (20, 64)):
        dxi, xi, u, it, kind = [], [], [], [], []
        for c in range(cases):
            mode = c % 4
            box = (0.5, 0.5) if mode == 1 else (1.5, 0.9)
            if N >= 10:
                box = (1.5, 0.9)
            x = np.stack([rng.uniform(-box[0], box[0], N), rng.uniform(-box[1], box[1], N)])
            if mode == 2:                          # an unsafe pair: h < 0 -> gain 1e6 branch
                i, j = rng.choice(N, 2, replace=False)
                ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(0.12, 0.2)
                x[:, j] = x[:, i] + d * np.array([np.cos(ang), np.sin(ang)])
            ang = rng.uniform(0, 2 * np.pi, N)
            mag = rng.uniform(0, 0.25 if mode == 3 else 0.15, N)      # > 0.2 exercises the pre-clip
            d = np.stack([mag * np.cos(ang), mag * np.sin(ang)])
            is_default = int(c % 8 == 7)
            (res), iters = _count_qp_iters(lambda: (default if is_default else safe)(d.copy(), x.copy()))
            dxi.append(d), xi.append(x), u.append(res), it.append(iters), kind.append(is_default)
        blob["N%d.dxi" % N], blob["N%d.xi" % N], blob["N%d.u" % N] = np.stack(dxi), np.stack(xi), np.stack(u)
        blob["N%d.iters" % N] = np.array(it, dtype=np.int32)
        blob["N%d.default" % N] = np.array(kind, dtype=np.int32)
        print("qp N=%2d  %4d cases  iters mean %.1f  min %d  max %d" % (N, cases, np.mean(it), min(it), max(it)))
    path = os.path.join(OUT, "qp_vectors.npz")
    np.savez_compressed(path, **blob)
    print("qp_vectors %.1f KB" % (os.path.getsize(path) / 1024.0))


PROJECTED = dict(rps_collision_offset=0.025)      # the heading-projected form of rps' collision test


def collision_variants():
    """The same kinds of cases under the other published form of rps' collision test (config keys
    rps_collision_offset / rps_collision_diameter; oracle/shims/rps/robotarium_abc.py)."""
    inject("PCP_projected_collision_inject", "PredatorCapturePrey", 192, seed=51, crowd=0.6, **PROJECTED)
    inject("Warehouse_projected_collision_inject", "Warehouse", 96, seed=52, crowd=0.6, **PROJECTED)
    rollout("MT_projected_collision_rollout", "MaterialTransport", 60, seed=53, **PROJECTED)
    inject("PCP20_projected_collision_inject", "PredatorCapturePrey", 24, seed=54, crowd=0.3, **dict(PCP20, **PROJECTED))


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["collision_variants"]:         # only the fixtures added in round 2 (the others are unchanged)
        return collision_variants()
    qp_vectors()
    collision_variants()
    for i, scn in enumerate(("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple")):
        rollout("%s_rollout" % scn, scn, 160, seed=11 + i)
        inject("%s_inject" % scn, scn, 256, seed=21 + i)
    # configuration variants of the primary scenario (+ MT capability_aware)
    inject("PCP_capability_aware_inject", "PredatorCapturePrey", 64, seed=31, capability_aware=True)
    inject("PCP_default_barrier_inject", "PredatorCapturePrey", 64, seed=32, barrier_certificate="default")
    inject("PCP_no_penalty_inject", "PredatorCapturePrey", 64, seed=33, penalize_violations=False)
    inject("PCP_robotarium_inject", "PredatorCapturePrey", 24, seed=34, robotarium=True)
    inject("PCP_neighbors2_inject", "PredatorCapturePrey", 64, seed=35, num_neighbors=2)
    inject("MT_capability_aware_inject", "MaterialTransport", 32, seed=36, capability_aware=True)
    inject("Warehouse_neighbors3_inject", "Warehouse", 64, seed=37, num_neighbors=3)
    rollout("PCP20_rollout", "PredatorCapturePrey", 40, seed=41, **PCP20)
    inject("PCP20_inject", "PredatorCapturePrey", 48, seed=42, crowd=0.2, **PCP20)


if __name__ == "__main__":
    main()
