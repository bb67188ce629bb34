"""Scenario YAML dict -> mrb_config.  Reads exactly the keys the reference's scenario classes read
(SURVEY.md Appendix B); spawn grids follow rps generate_initial_conditions + utilities/misc.py:49-63."""
import math
import os

import yaml

from . import _lib

SCENARIOS = ("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple")
GYM_KEYS = {s: s + "-v0" for s in SCENARIOS}                      # robotarium_gym/__init__.py:4-10
CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")
# defaults of the two rps collision constants (overridable per config with rps_collision_diameter / rps_collision_offset,
# or for a whole process with the environment variables below)
RPS_COLLISION_DIAMETER = float(os.environ.get("MRB_RPS_COLLISION_DIAMETER", "0.135"))
RPS_COLLISION_OFFSET = float(os.environ.get("MRB_RPS_COLLISION_OFFSET", "0.0"))


class objectview(object):
    """utilities/misc.py:44-47: attribute view of the YAML dict."""

    def __init__(self, d):
        self.__dict__ = d


def default_config_path(scenario):
    return os.path.join(CONFIG_DIR, scenario + ".yaml")


def load_yaml(path):
    with open(path, "r") as f:
        return yaml.safe_load(f)


def _grid(count, spacing, width, height, sx1=0.0, sx2=0.0, sy1=0.0, sy2=0.0, random_theta=0):
    xr, yr = int(math.floor(width / spacing)), int(math.floor(height / spacing))
    if xr == 0 or yr == 0 or xr * yr <= count:            # rps' assert
        raise ValueError("Cannot fit %d items on a %dx%d spawn grid (spacing %g)" % (count, xr, yr, spacing))
    return _lib.Spawn(count, xr, yr, random_theta, spacing, width / 2, height / 2, sx1, sx2, sy1, sy2)


def _locations(count, width, height, thresh, start_dist, spawn_left=True):
    shift = width / 2 - thresh
    return _grid(count, start_dist, width, height, sx1=-shift if spawn_left else shift)


def make_config(scenario, cfg, auto_reset=False, track_dist=True, collect_stats=True):
    if scenario not in SCENARIOS:
        raise KeyError(scenario)
    g = cfg.get
    c = _lib.Config()
    c.struct_size = _lib.C.sizeof(_lib.Config)
    c.scenario = SCENARIOS.index(scenario)
    c.update_frequency = int(cfg["update_frequency"])
    c.ctrl_period = 15                                             # roboEnv.py:63
    c.robotarium = int(bool(g("robotarium", False)))
    c.penalize_violations = int(bool(g("penalize_violations", True)))
    kind = g("barrier_certificate", "safe")                        # roboEnv.py:15-18: absent -> 'safe'
    if kind not in ("safe", "default"):
        raise ValueError("barrier_certificate must be 'safe' or 'default' (utilities/controller.py:13-18)")
    c.barrier_default = int(kind == "default")
    c.max_episode_steps = int(cfg["max_episode_steps"])
    c.num_neighbors = int(g("num_neighbors", 0))
    c.capability_aware = int(bool(g("capability_aware", False)))
    c.auto_reset, c.track_dist, c.collect_stats = int(auto_reset), int(track_dist), int(collect_stats)
    c.left, c.right, c.up, c.down = cfg["LEFT"], cfg["RIGHT"], cfg["UP"], cfg["DOWN"]
    # rps RobotariumABC._validate's collision test (not a key of the reference's config.yaml: rps hard-codes it).
    # rps is un-vendored and its pinned commit cannot be read here, so both published forms are selectable:
    # centre to centre (offset 0, the restatement every fixture was made with) or heading-projected points (0.025)
    c.collision_diameter = float(g("rps_collision_diameter", RPS_COLLISION_DIAMETER))
    c.collision_offset = float(g("rps_collision_offset", RPS_COLLISION_OFFSET))
    height = cfg["DOWN"] - cfg["UP"]
    if scenario == "PredatorCapturePrey":
        c.num_robots = cfg["predator"] + cfg["capture"]
        c.num_prey, c.num_predators = cfg["num_prey"], cfg["predator"]
        c.predator_radius, c.capture_radius = cfg["predator_radius"], cfg["capture_radius"]
        c.step_dist = cfg["step_dist"]
        c.time_penalty, c.sense_reward, c.capture_reward = cfg["time_penalty"], cfg["sense_reward"], cfg["capture_reward"]
        c.violation_reward = -5                                    # PredatorCapturePrey.py:159
        c.spawn_robots = _locations(c.num_robots, cfg["ROBOT_INIT_RIGHT_THRESH"] - cfg["LEFT"], height,
                                    cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["start_dist"])
        c.spawn_other = _locations(c.num_prey, cfg["RIGHT"] - cfg["PREY_INIT_LEFT_THRESH"], height,
                                   cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["step_dist"], spawn_left=False)
    elif scenario == "Warehouse":
        c.num_robots = cfg["n_agents"]
        c.step_dist = cfg["step_dist"]
        c.load_reward, c.unload_reward, c.goal_width = cfg["load_reward"], cfg["unload_reward"], cfg["goal_width"]
        c.violation_reward = -5                                    # warehouse.py:116
        c.spawn_robots = _grid(c.num_robots, cfg["start_dist"], cfg["RIGHT"] - cfg["LEFT"], height,   # warehouse.py:91-98
                               sx1=(1.5 + cfg["LEFT"]) / 2, sx2=-((1.5 - cfg["RIGHT"]) / 2),
                               sy1=-((1 + cfg["UP"]) / 2), sy2=(1 - cfg["DOWN"]) / 2, random_theta=1)
    elif scenario == "MaterialTransport":
        c.num_robots = cfg["n_agents"]
        c.n_fast = cfg["n_fast_agents"]
        if cfg["n_fast_agents"] + cfg["n_slow_agents"] != c.num_robots:
            raise ValueError("n_fast_agents + n_slow_agents must equal n_agents")
        c.small_torque, c.large_torque = cfg["small_torque"], cfg["large_torque"]
        c.fast_step, c.slow_step = cfg["fast_step"], cfg["slow_step"]
        c.time_penalty = cfg["time_penalty"]
        c.load_reward, c.unload_reward, c.goal_width = cfg["load_multiplier"], cfg["unload_multiplier"], cfg["end_goal_width"]
        c.zone1_radius = cfg["zone1_radius"]
        c.violation_reward = -6                                    # MaterialTransport.py:137
        for k, z in enumerate(("zone1", "zone2")):
            if cfg[z]["distribution"] != "normal":
                raise ValueError("only the 'normal' zone load distribution of the shipped config is supported")
            c.zone_mu[k], c.zone_sigma[k] = cfg[z]["loc"], cfg[z]["scale"]
        c.spawn_robots = _locations(c.num_robots, cfg["end_goal_width"], height,
                                    cfg["LEFT"] + cfg["end_goal_width"], cfg["start_dist"])
    elif scenario == "ArcticTransport":
        c.num_robots = cfg["n_agents"]
        c.step_dist, c.fast_step, c.slow_step = cfg["normal_step"], cfg["fast_step"], cfg["slow_step"]
        c.not_reached_penalty, c.dist_multiplier = cfg["not_reached_penalty"], cfg["dist_multiplier"]
        c.violation_reward = -30                                   # ArcticTransport.py:103
    else:
        c.num_robots = cfg["n_agents"]
        c.step_dist, c.reward_scaler = cfg["step_dist"], cfg["reward_scaler"]
        c.violation_reward = -5                                    # simple.py:174
        c.spawn_robots = _locations(c.num_robots, cfg["ROBOT_INIT_RIGHT_THRESH"] - cfg["LEFT"], height,
                                    cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["start_dist"])
        c.spawn_other = _locations(1, cfg["RIGHT"] - cfg["PREY_INIT_LEFT_THRESH"], height,
                                   cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["step_dist"], spawn_left=False)
    return c


def obs_space_dim(scenario, cfg, c):
    """Declared observation_space width (the reference declares agent_obs_dim*(num_neighbors+1) even
    when fewer neighbours exist, e.g. PredatorCapturePrey.py:52)."""
    if scenario == "PredatorCapturePrey":
        return (6 if c.capability_aware else 4) * (c.num_neighbors + 1)
    if scenario == "Warehouse":
        return 3 * (c.num_neighbors + 1)
    if scenario == "MaterialTransport":
        return 11 if c.capability_aware else 9
    if scenario == "ArcticTransport":
        return 30
    return 2 * (c.num_robots + 1)
