"""Run /root/reference/robotarium_gym UNMODIFIED on the stand-in packages in oracle/shims.

ORACLE / TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container); the GPU box never imports this file.  Used by oracle/gen_golden.py to produce the
committed fixtures under tests/golden/ and by the `not gpu` tests that pin the C restatement
(oracle/marbler_oracle.c) to the reference's own scenario / roboEnv / Controller code.

State injection follows SURVEY.md section 8c workaround (3): overwrite scn.agent_poses IN PLACE
(it aliases the simulator state), set the scenario flags, set roboEnv.previous_pose and
re-synchronise roboEnv.errors with the simulator's cumulative error dict.
"""
import contextlib
import copy
import io
import os
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
REFERENCE_ROOT = os.environ.get("MARBLER_REFERENCE", "/root/reference")

SCENARIOS = ("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple")
MSG_CODE = {"": 0, "collision": 1, "boundary": 2, "collision_boundary": 3}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "robotarium_gym"))


def _ensure_paths():
    for p in (REFERENCE_ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)


def reference_config(scenario, **overrides):
    """The scenario's own config.yaml with figures / logging / gif off."""
    path = os.path.join(REFERENCE_ROOT, "robotarium_gym", "scenarios", scenario, "config.yaml")
    with open(path) as f:
        cfg = yaml.safe_load(f)
    cfg.update(dict(show_figure_frequency=-1, save_gif=False, enable_logging=False,
                    real_time=False, robotarium=False))
    cfg.update(overrides)
    return cfg


class RefEnv(object):
    """One reference env (robotarium_gym.wrapper.Wrapper) with inject / extract helpers."""

    def __init__(self, scenario, **overrides):
        _ensure_paths()
        import rps.robotarium_abc as rabc
        # workaround (2): rps' cumulative error dict is process global; start every env clean
        rabc.RobotariumABC._validate.__defaults__[0].clear()
        from robotarium_gym.wrapper import Wrapper
        self.scenario = scenario
        self.cfg = reference_config(scenario, **overrides)
        fd, self._cfg_path = tempfile.mkstemp(suffix=".yaml")
        with os.fdopen(fd, "w") as f:
            yaml.safe_dump(self.cfg, f)
        self.wrapper = Wrapper(scenario, self._cfg_path)
        self.scn = self.wrapper.env
        self.N = self.scn.num_robots
        self._sink = io.StringIO()

    def __del__(self):
        try:
            os.unlink(self._cfg_path)
        except Exception:
            pass

    # ------------------------------------------------------------------ reset / step
    def reset(self):
        # the rps collision test is not a MARBLER config key (rps hard-codes it); two extra keys select the form the
        # stand-in applies.  roboEnv builds a new Robotarium at every reset (roboEnv.py:98-112), which reads these.
        import rps.robotarium_abc as rabc
        rabc.COLLISION_OFFSET = float(self.cfg.get("rps_collision_offset", 0.0))
        rabc.COLLISION_DIAMETER = float(self.cfg.get("rps_collision_diameter", 0.135))
        with contextlib.redirect_stdout(self._sink):
            obs = self.wrapper.reset()
        self.scn.env.errors = copy.deepcopy(self.scn.env.robotarium._errors)
        return obs

    def step(self, actions):
        if self.scenario == "PredatorCapturePrey":
            self.scn.prey_locs = []          # workaround (1): numpy-2 crash at PredatorCapturePrey.py:185
        with contextlib.redirect_stdout(self._sink):
            obs, rew, done, info = self.wrapper.step([int(a) for a in actions])
        self._sink.seek(0)
        self._sink.truncate(0)
        msg = info.get("message", "")
        if self.scenario == "Simple" and "remaining" in info:   # simple.py:176 files the message here
            msg = info["remaining"]
        out = {
            "obs": np.array([np.asarray(o, dtype=np.float64) for o in obs]),
            "reward": np.asarray(rew, dtype=np.float64),
            "done": np.asarray(done, dtype=np.bool_),
            "message": np.int32(MSG_CODE[msg]),
            "dist": np.asarray(info["dist_travelled"], dtype=np.float64),
        }
        return out

    # ------------------------------------------------------------------ state
    def get_state(self):
        scn = self.scn
        st = {"poses": np.array(scn.agent_poses, dtype=np.float64),
              "episode_steps": np.int32(scn.episode_steps)}
        pp = scn.env.previous_pose
        st["prev_valid"] = np.int32(pp is not None)
        st["prev_pose"] = np.zeros((3, self.N)) if pp is None else np.array(pp, dtype=np.float64)
        s = self.scenario
        if s == "PredatorCapturePrey":
            st["prey_loc"] = np.array(scn.prey_loc, dtype=np.float64)
            st["prey_sensed"] = np.array(scn.prey_sensed, dtype=np.uint8)
            st["prey_captured"] = np.array(scn.prey_captured, dtype=np.uint8)
        elif s == "Warehouse":
            st["loaded"] = np.array([a.loaded for a in scn.agents], dtype=np.uint8)
        elif s == "MaterialTransport":
            st["load"] = np.array([a.load for a in scn.agents], dtype=np.int32)
            st["zone_load"] = np.array([scn.zone1_load, scn.zone2_load], dtype=np.int32)
            st["messages"] = np.array(scn.messages, dtype=np.int32)
        elif s == "ArcticTransport":
            st["grid"] = np.array(scn.grid, dtype=np.uint8)
            st["goal_col"] = np.int32(scn.goal_loc[1])
            st["pixel_type"] = np.array([a.pixel_type for a in scn.agents], dtype=np.int32)
            st["reached_goal"] = np.array([a.reached_goal for a in scn.agents], dtype=np.uint8)
        elif s == "Simple":
            st["goal"] = np.array(scn.goal_loc, dtype=np.float64).reshape(2)
        return st

    def set_state(self, st):
        scn = self.scn
        scn.agent_poses[...] = st["poses"]                 # in place: aliases robotarium.poses
        assert scn.agent_poses is scn.env.robotarium.poses
        scn.episode_steps = int(st["episode_steps"])
        scn.env.previous_pose = np.array(st["prev_pose"], dtype=np.float64) if int(st["prev_valid"]) else None
        scn.env.errors = copy.deepcopy(scn.env.robotarium._errors)
        s = self.scenario
        if s == "PredatorCapturePrey":
            scn.prey_loc = np.array(st["prey_loc"], dtype=np.float64)
            scn.prey_sensed = [bool(v) for v in st["prey_sensed"]]
            scn.prey_captured = [bool(v) for v in st["prey_captured"]]
            scn.state_space = scn._generate_state_space()
        elif s == "Warehouse":
            for a, v in zip(scn.agents, st["loaded"]):
                a.loaded = bool(v)
        elif s == "MaterialTransport":
            for a, v in zip(scn.agents, st["load"]):
                a.load = int(v)
            scn.zone1_load, scn.zone2_load = int(st["zone_load"][0]), int(st["zone_load"][1])
            scn.messages = [int(v) for v in st["messages"]]
        elif s == "ArcticTransport":
            scn.grid = np.array(st["grid"], dtype=np.int64)
            scn.goal_loc = [1, int(st["goal_col"])]
            for a, p, r in zip(scn.agents, st["pixel_type"], st["reached_goal"]):
                a.pixel_type = int(p)
                a.reached_goal = bool(r)
        elif s == "Simple":
            scn.goal_loc = np.array(st["goal"], dtype=np.float64).reshape(1, 2)
            scn.state_space = scn._generate_state_space()

    def step_from(self, st, actions):
        """Inject `st`, take one step, return (outputs, post-state)."""
        self.set_state(st)
        out = self.step(actions)
        return out, self.get_state()


def last_qp_iterations():
    _ensure_paths()
    import cvxopt.solvers as s
    return s.last["iterations"]
