"""Loader for tests/golden/*.npz (made by oracle/gen_golden.py from the reference run on oracle/shims)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DISCRETE_STATE = ("episode_steps", "prev_valid", "prey_sensed", "prey_captured", "loaded", "load", "zone_load",
                  "messages", "grid", "goal_col", "pixel_type", "reached_goal")


def fixture_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not p.endswith("qp_vectors.npz"))


class Golden(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.scenario = str(z["scenario"])
        self.cfg = json.loads(str(z["cfg_json"]))
        self.actions = z["actions"]
        self.qp_iters = z["qp_iters"]
        self.s0 = {k[3:]: z[k] for k in z.files if k.startswith("s0.")}
        self.s1 = {k[3:]: z[k] for k in z.files if k.startswith("s1.")}
        self.out = {k[4:]: z[k] for k in z.files if k.startswith("out.")}
        self.B = self.actions.shape[0]


def qp_vectors():
    z = np.load(os.path.join(GOLDEN, "qp_vectors.npz"))
    out = {}
    for k in z.files:
        n, f = k.split(".")
        out.setdefault(int(n[1:]), {})[f] = z[k]
    return out


def compare_step(g, out, s1, pose_tol, obs_tol, rew_tol, dist_tol, skip=()):
    """Assert a batched (out, s1) pair against the fixture.  Discrete fields must be bit-exact."""
    assert np.array_equal(np.asarray(out["message"]).astype(np.int64), g.out["message"].astype(np.int64)), "message"
    done = np.asarray(out["done"])
    done = done if done.ndim == 1 else done[:, 0]
    assert np.array_equal(done.astype(bool), g.out["done"][:, 0]), "done"
    for k in DISCRETE_STATE:
        if k in g.s1 and k not in skip:
            assert np.array_equal(np.asarray(s1[k]).astype(np.int64).reshape(g.s1[k].shape),
                                  g.s1[k].astype(np.int64)), k
    err = {
        "poses": np.abs(np.asarray(s1["poses"]) - g.s1["poses"]).max(),
        "obs": np.abs(np.asarray(out["obs"], dtype=np.float64) - g.out["obs"]).max(),
        "reward": np.abs(np.asarray(out["reward"], dtype=np.float64) - g.out["reward"]).max(),
        "dist": np.abs(np.asarray(out["dist"], dtype=np.float64) - g.out["dist"]).max(),
    }
    # theta compares modulo 2*pi (atan2 wrap at +-pi)
    dth = np.asarray(s1["poses"])[:, 2] - g.s1["poses"][:, 2]
    dth = np.abs(np.arctan2(np.sin(dth), np.cos(dth))).max()
    err["poses"] = max(np.abs(np.asarray(s1["poses"])[:, :2] - g.s1["poses"][:, :2]).max(), dth)
    assert err["poses"] <= pose_tol, err
    assert err["obs"] <= obs_tol, err
    assert err["reward"] <= rew_tol, err
    assert err["dist"] <= dist_tol, err
    return err


EPISODES = os.path.join(GOLDEN, "episodes")


def episode_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(EPISODES, "*.npz")))


class Episodes(object):
    """Whole reference trajectories from a config seed (oracle/gen_episodes.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(EPISODES, name + ".npz"))
        self.scenario = str(z["scenario"])
        self.cfg = json.loads(str(z["cfg_json"]))
        self.py_seed = int(z["py_seed"])
        for k in ("actions", "obs", "reward", "done", "message", "dist", "qp_max_iters", "reset_after"):
            setattr(self, k, z[k])
        self.resets = {k[6:]: z[k] for k in z.files if k.startswith("reset.")}
        self.T = len(self.actions)
