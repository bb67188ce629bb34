"""Seed-for-seed episodes: with `num_envs == 1` a reset draws from numpy's global
legacy RNG (and Python's random) exactly as the reference does, so whole multi-episode trajectories follow
the reference's for the same config seed.  Fixtures: tests/golden/episodes (oracle/gen_episodes.py, the
unmodified reference on oracle/shims)."""
import random

import numpy as np
import pytest

import golden_util as gu

STALL = 25
TRAJ_TOL = 1e-5 + 2e-6


def _budget(ep, compared, lost):
    """Whole trajectories are compared step by step at the 1e-5 bar.  cvxopt's stopping rule with rps' loose
    tolerances (1e-2) is a discontinuity: when a residual sits within rounding of its threshold two
    implementations stop one iteration apart and the returned velocities differ by ~1e-3, after which the
    trajectories separate for the rest of that episode.  That may happen at most once per fixture (150 steps,
    ~300 solves) and at least 90 % of all steps must have been compared in sync.  Measured (C oracle and CUDA path
    alike): nine of the ten fixtures are followed for all 150 steps, MaterialTransport_seed11 loses the last 10 steps
    of one episode."""
    print("episode sync: %d of %d steps compared, %d loss(es) of sync" % (compared, ep.T, lost))
    assert lost <= 1, lost
    assert compared >= 0.9 * ep.T, (compared, ep.T)


def _same_state(a, b, env_axis=None):
    for k, v in a.items():
        w = np.asarray(b[k]) if env_axis is None else np.asarray(b[k])[env_axis]
        assert np.array_equal(np.asarray(v), w.reshape(np.asarray(v).shape)), k


@pytest.mark.parametrize("name", gu.episode_names())
def test_reset_sampler_reproduces_reference_draws(name):
    """Pure host logic: the sampler consumes numpy's global stream call for call like the reference."""
    from marbler_b200.reference_rng import sample_reset
    ep = gu.Episodes(name)
    np.random.seed(ep.cfg["seed"])
    random.seed(ep.py_seed)
    n = len(ep.resets["poses"])
    for e in range(n):
        st = sample_reset(ep.scenario, ep.cfg)
        _same_state({k: v[e] for k, v in ep.resets.items()}, st)


@pytest.mark.parametrize("name", gu.episode_names())
def test_oracle_follows_reference_episodes(oracle_lib, name):
    """The C restatement started from the sampler's states tracks the reference over whole episodes."""
    from marbler_b200.reference_rng import sample_reset
    ep = gu.Episodes(name)
    orc = oracle_lib.COracle(ep.scenario, ep.cfg)
    np.random.seed(ep.cfg["seed"])
    random.seed(ep.py_seed)
    st = sample_reset(ep.scenario, ep.cfg)
    synced, compared, lost = True, 0, 0
    for t in range(ep.T):
        out, st = orc.step(st, ep.actions[t])
        if ep.qp_max_iters[t] >= STALL:
            synced = False                       # limit-cycle solve: not reproducible
        if synced:
            ok = int(out["message"]) == ep.message[t] and bool(out["done"][0]) == bool(ep.done[t]) and \
                max(np.abs(out["obs"] - ep.obs[t]).max(), np.abs(out["reward"] - ep.reward[t]).max(),
                    np.abs(out["dist"] - ep.dist[t]).max()) < TRAJ_TOL
            if ok:
                compared += 1
            else:                                # a borderline stopping test flipped: see _budget()
                synced, lost = False, lost + 1
        if ep.reset_after[t]:
            st = sample_reset(ep.scenario, ep.cfg)
            synced = True
    _budget(ep, compared, lost)


@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.episode_names())
def test_wrapper_follows_reference_episodes(name, tmp_path):
    """The CUDA path behind the reference's Wrapper surface: same seed -> same episodes as the reference."""
    import yaml
    import marbler_b200
    ep = gu.Episodes(name)
    path = tmp_path / "config.yaml"
    path.write_text(yaml.safe_dump(ep.cfg))
    env = marbler_b200.Wrapper(ep.scenario, str(path))            # num_envs = 1: reset_rng = "reference"
    random.seed(ep.py_seed)
    obs = env.reset()
    assert not np.any(np.asarray(obs))
    _same_state({k: v[0] for k, v in ep.resets.items() if k != "prev_pose"}, env.env.get_state(), env_axis=0)
    from marbler_b200.scenarios.base import MESSAGES
    synced, episode, compared, lost = True, 0, 0, 0
    for t in range(ep.T):
        obs, rew, done, info = env.step(list(ep.actions[t]))
        if ep.qp_max_iters[t] >= STALL:
            synced = False
        if synced:
            msg = info.get("message", "")
            if ep.scenario == "Simple" and "remaining" in info:
                msg = info["remaining"]
            ok = MESSAGES.index(msg) == ep.message[t] and done == [bool(ep.done[t])] * env.n_agents and \
                max(np.abs(np.asarray(obs, dtype=np.float64) - ep.obs[t]).max(),
                    np.abs(np.asarray(rew) - ep.reward[t]).max(),
                    np.abs(info["dist_travelled"] - ep.dist[t]).max()) < TRAJ_TOL
            if ok:
                compared += 1
            else:
                synced, lost = False, lost + 1
        if ep.reset_after[t]:
            env.reset()
            episode += 1
            synced = True
            _same_state({k: v[episode] for k, v in ep.resets.items() if k != "prev_pose"}, env.env.get_state(), env_axis=0)
    assert episode == len(ep.resets["poses"]) - 1
    _budget(ep, compared, lost)
