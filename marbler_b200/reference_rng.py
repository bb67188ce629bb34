"""Seed-for-seed reset for the single-env (`num_envs == 1`) drop-in path  (SURVEY.md section 8f-3).

The reference resets an episode by drawing from numpy's GLOBAL legacy RandomState (seeded once in the
scenario constructor when `seed != -1`, e.g. PredatorCapturePrey.py:27-28) and - ArcticTransport only - from
Python's `random` module (ArcticTransport.py:72, never seeded).  The batched kernels use a counter-based
Philox stream instead (one stream per env, reproducible under sharding), which has the same distribution but
not the same numbers.  For `num_envs == 1` this module reproduces the reference's own draw sequence, call for
call, on the host with numpy itself and hands the result to the device with `set_state`; whole episodes then
start from exactly the state the reference would start from for the same seed.

Draw order per scenario (file:line under robotarium_gym/):
  PredatorCapturePrey.py:112-132   robots: choice(cells, N, replace=False), then one rand() per robot (the
                                   heading rps draws and utilities/misc.py:57 overwrites with 0); prey: the same
  warehouse.py:86-98               choice + rand per robot; headings kept, then the four net shifts
  MaterialTransport.py:95-110      int(normal(**zone1)), int(normal(**zone2)), then robots
  ArcticTransport.py:56-78         np.random.randint(3, size=(8, 12)), random.randint(1, 11)
  simple.py:129-147                robots, then the goal (one location on the prey grid)
followed by roboEnv.reset -> _create_robotarium -> one zero-velocity Robotarium.step (roboEnv.py:108-112),
which only wraps the headings: theta = arctan2(sin theta, cos theta).
"""
import math
import random as _pyrandom

import numpy as np


def _initial_conditions(rng, n, spacing, width, height):
    """rps.utilities.misc.generate_initial_conditions."""
    xr, yr = int(np.floor(width / spacing)), int(np.floor(height / spacing))
    assert xr != 0 and yr != 0
    assert xr * yr > n, "Cannot fit %d robots on a %dx%d grid" % (n, xr, yr)
    choices = rng.choice(xr * yr, n, replace=False)
    poses = np.zeros((3, n))
    for i, c in enumerate(choices):
        x, y = divmod(c, yr)
        poses[0, i] = x * spacing - width / 2
        poses[1, i] = y * spacing - height / 2
        poses[2, i] = rng.rand() * 2 * np.pi - np.pi
    return poses


def _initial_locations(rng, n, width, height, thresh, start_dist=.3, spawn_left=True):
    """utilities/misc.py:49-63 generate_initial_locations."""
    poses = _initial_conditions(rng, n, start_dist, width, height)
    for i in range(n):
        if spawn_left:
            poses[0][i] -= (width / 2 - thresh)
        else:
            poses[0][i] += (width / 2 - thresh)
        poses[2][i] = 0
    return poses


def sample_reset(scenario, cfg, rng=None, pyrandom=None):
    """One initial state (dict of arrays WITHOUT a leading env axis, keys of marbler_b200.layout) drawn with
    the reference's sequence of RNG calls.  rng: a numpy RandomState-like (default: the global np.random, as
    the reference); pyrandom: a `random`-module-like (default: the global module)."""
    rng = np.random if rng is None else rng
    pyrandom = _pyrandom if pyrandom is None else pyrandom
    g = cfg.get
    st = {"episode_steps": np.int32(0), "prev_valid": np.int32(0)}
    if scenario == "PredatorCapturePrey":
        n = cfg["predator"] + cfg["capture"]
        height = cfg["DOWN"] - cfg["UP"]
        width = cfg["ROBOT_INIT_RIGHT_THRESH"] - cfg["LEFT"]
        poses = _initial_locations(rng, n, width, height, cfg["ROBOT_INIT_RIGHT_THRESH"], start_dist=cfg["start_dist"])
        width = cfg["RIGHT"] - cfg["PREY_INIT_LEFT_THRESH"]
        prey = _initial_locations(rng, cfg["num_prey"], width, height, cfg["ROBOT_INIT_RIGHT_THRESH"],
                                  start_dist=cfg["step_dist"], spawn_left=False)
        st["prey_loc"] = prey[:2].T.copy()
        st["prey_sensed"] = np.zeros(cfg["num_prey"], dtype=np.uint8)
        st["prey_captured"] = np.zeros(cfg["num_prey"], dtype=np.uint8)
    elif scenario == "Warehouse":
        n = cfg["n_agents"]
        width, height = cfg["RIGHT"] - cfg["LEFT"], cfg["DOWN"] - cfg["UP"]
        poses = _initial_conditions(rng, n, cfg["start_dist"], width, height)
        poses[0] += (1.5 + cfg["LEFT"]) / 2
        poses[0] -= (1.5 - cfg["RIGHT"]) / 2
        poses[1] -= (1 + cfg["UP"]) / 2
        poses[1] += (1 - cfg["DOWN"]) / 2
        st["loaded"] = np.zeros(n, dtype=np.uint8)
    elif scenario == "MaterialTransport":
        n = cfg["n_agents"]
        zones = []
        for z in ("zone1", "zone2"):
            args = dict(cfg[z])
            dist = args.pop("distribution")
            zones.append(int(getattr(rng, dist)(**args)))
        poses = _initial_locations(rng, n, cfg["end_goal_width"], cfg["DOWN"] - cfg["UP"],
                                   cfg["LEFT"] + cfg["end_goal_width"], start_dist=cfg["start_dist"])
        st["load"] = np.zeros(n, dtype=np.int32)
        st["zone_load"] = np.array(zones, dtype=np.int32)
        st["messages"] = np.zeros(4, dtype=np.int32)
    elif scenario == "ArcticTransport":
        n = cfg["n_agents"]
        poses = np.array([[-.3, .3, -.9, .9], [-.8] * 4, [math.pi / 2] * n])
        grid = rng.randint(3, size=(8, 12))
        goal = pyrandom.randint(1, 11)
        grid[0][goal] = 3
        grid[0][goal - 1] = 3
        grid[1][goal] = 3
        grid[1][goal - 1] = 3
        grid[7][1:11] = 0
        st["grid"] = grid.astype(np.uint8)
        st["goal_col"] = np.int32(goal)
        st["pixel_type"] = np.zeros(n, dtype=np.int32)
        st["reached_goal"] = np.zeros(n, dtype=np.uint8)
    elif scenario == "Simple":
        n = cfg["n_agents"]
        height = cfg["DOWN"] - cfg["UP"]
        width = cfg["ROBOT_INIT_RIGHT_THRESH"] - cfg["LEFT"]
        poses = _initial_locations(rng, n, width, height, cfg["ROBOT_INIT_RIGHT_THRESH"], start_dist=cfg["start_dist"])
        width = cfg["RIGHT"] - cfg["PREY_INIT_LEFT_THRESH"]
        goal = _initial_locations(rng, 1, width, height, cfg["ROBOT_INIT_RIGHT_THRESH"],
                                  start_dist=cfg["step_dist"], spawn_left=False)
        st["goal"] = goal[:2].T.reshape(2).copy()
    else:
        raise KeyError(scenario)
    del g
    # the zero-velocity simulator step of roboEnv._create_robotarium (roboEnv.py:111-112): x, y unchanged
    poses[2] = np.arctan2(np.sin(poses[2]), np.cos(poses[2]))
    st["poses"] = poses
    st["prev_pose"] = np.zeros_like(poses)
    return st
