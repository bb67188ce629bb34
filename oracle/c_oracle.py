"""ctypes view of oracle/marbler_oracle.c.  ORACLE / TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under marbler_b200/ does.  State travels as the same dict-of-arrays that
oracle/ref_harness.py's RefEnv.get_state()/set_state() use, with an optional leading batch axis.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libmarbler_oracle.so")
SCENARIOS = ("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple")
N_ACTIONS = {"PredatorCapturePrey": 5, "Warehouse": 5, "MaterialTransport": 20, "ArcticTransport": 5, "Simple": 5}


class Spawn(C.Structure):
    _fields_ = [("count", C.c_int32), ("xr", C.c_int32), ("yr", C.c_int32), ("random_theta", C.c_int32),
                ("spacing", C.c_double), ("w2", C.c_double), ("h2", C.c_double),
                ("sx1", C.c_double), ("sx2", C.c_double), ("sy1", C.c_double), ("sy2", C.c_double)]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "scenario", "N", "update_frequency", "ctrl_period", "robotarium", "penalize_violations",
        "barrier_default", "max_episode_steps", "num_neighbors", "capability_aware", "num_prey",
        "num_predators", "n_fast", "small_torque", "large_torque", "zone1_loc_unused")] + \
        [(n, C.c_double) for n in (
            "LEFT", "RIGHT", "UP", "DOWN", "step_dist", "fast_step", "slow_step", "predator_radius",
            "capture_radius", "time_penalty", "sense_reward", "capture_reward", "load_reward",
            "unload_reward", "goal_width", "zone1_radius", "not_reached_penalty", "dist_multiplier",
            "reward_scaler", "violation_reward")] + \
        [("zone_mu", C.c_double * 2), ("zone_sigma", C.c_double * 2),
         ("spawn_robots", Spawn), ("spawn_other", Spawn),
         ("collision_diameter", C.c_double), ("collision_offset", C.c_double)]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "marbler_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        cp = C.POINTER(Config)
        for f in (L.orc_nf, L.orc_ni, L.orc_obs_dim):
            f.argtypes, f.restype = [cp], C.c_int
        L.orc_step_batch.argtypes = [cp, C.c_int64, dp, ip, ip, dp, dp, dp, ip, C.c_int, C.c_uint64, C.c_uint64]
        L.orc_step_batch.restype = None
        L.orc_reset_batch.argtypes = [cp, C.c_int64, C.c_uint64, C.c_uint64, dp, ip]
        L.orc_reset_batch.restype = None
        L.orc_barrier_qp.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_barrier_qp.restype = C.c_int
        L.orc_controller.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_controller.restype = C.c_int
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_philox4x32_10.restype = None
        _lib = L
    return _lib


def _grid(count, spacing, width, height, sx1=0.0, sx2=0.0, sy1=0.0, sy2=0.0, random_theta=0):
    """rps generate_initial_conditions' grid with the shifts MARBLER applies."""
    xr, yr = int(np.floor(width / spacing)), int(np.floor(height / spacing))
    assert xr * yr > count, "Cannot fit %d robots on a %dx%d grid" % (count, xr, yr)
    return Spawn(count, xr, yr, random_theta, spacing, width / 2, height / 2, sx1, sx2, sy1, sy2)


def _locations(count, width, height, thresh, start_dist, spawn_left=True):
    """utilities/misc.py:49-63 generate_initial_locations."""
    shift = (width / 2 - thresh)
    return _grid(count, start_dist, width, height, sx1=-shift if spawn_left else shift)


def make_config(scenario, cfg):
    """Scenario name + the reference's YAML dict -> orc_config (reads the same keys the reference reads)."""
    g = cfg.get
    c = Config()
    c.scenario = SCENARIOS.index(scenario)
    c.update_frequency = cfg["update_frequency"]
    c.ctrl_period = 15                                        # roboEnv.py:63
    c.robotarium = int(bool(g("robotarium", False)))
    c.penalize_violations = int(bool(g("penalize_violations", True)))
    c.barrier_default = {"safe": 0, "default": 1}[g("barrier_certificate", "safe")]   # roboEnv.py:15-18
    c.max_episode_steps = cfg["max_episode_steps"]
    c.num_neighbors = g("num_neighbors", 0)
    c.capability_aware = int(bool(g("capability_aware", False)))
    c.LEFT, c.RIGHT, c.UP, c.DOWN = cfg["LEFT"], cfg["RIGHT"], cfg["UP"], cfg["DOWN"]
    # the rps collision test in force (oracle/shims/rps/robotarium_abc.py; ref_harness applies the same two keys)
    c.collision_diameter = float(g("rps_collision_diameter", 0.135))
    c.collision_offset = float(g("rps_collision_offset", 0.0))
    height = cfg["DOWN"] - cfg["UP"]
    if scenario == "PredatorCapturePrey":
        c.N = cfg["predator"] + cfg["capture"]                # PredatorCapturePrey.py:19
        c.num_prey, c.num_predators = cfg["num_prey"], cfg["predator"]
        c.predator_radius, c.capture_radius = cfg["predator_radius"], cfg["capture_radius"]
        c.step_dist = cfg["step_dist"]
        c.time_penalty, c.sense_reward, c.capture_reward = cfg["time_penalty"], cfg["sense_reward"], cfg["capture_reward"]
        c.violation_reward = -5
        c.spawn_robots = _locations(c.N, cfg["ROBOT_INIT_RIGHT_THRESH"] - cfg["LEFT"], height,
                                    cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["start_dist"])
        c.spawn_other = _locations(c.num_prey, cfg["RIGHT"] - cfg["PREY_INIT_LEFT_THRESH"], height,
                                   cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["step_dist"], spawn_left=False)
    elif scenario == "Warehouse":
        c.N = cfg["n_agents"]
        c.step_dist = cfg["step_dist"]
        c.load_reward, c.unload_reward, c.goal_width = cfg["load_reward"], cfg["unload_reward"], cfg["goal_width"]
        c.violation_reward = -5
        c.spawn_robots = _grid(c.N, cfg["start_dist"], cfg["RIGHT"] - cfg["LEFT"], height,      # warehouse.py:91-98
                               sx1=(1.5 + cfg["LEFT"]) / 2, sx2=-((1.5 - cfg["RIGHT"]) / 2),
                               sy1=-((1 + cfg["UP"]) / 2), sy2=(1 - cfg["DOWN"]) / 2, random_theta=1)
    elif scenario == "MaterialTransport":
        c.N = cfg["n_agents"]
        c.n_fast = cfg["n_fast_agents"]
        c.small_torque, c.large_torque = cfg["small_torque"], cfg["large_torque"]
        c.fast_step, c.slow_step = cfg["fast_step"], cfg["slow_step"]
        c.time_penalty = cfg["time_penalty"]
        c.load_reward, c.unload_reward, c.goal_width = cfg["load_multiplier"], cfg["unload_multiplier"], cfg["end_goal_width"]
        c.zone1_radius = cfg["zone1_radius"]
        c.violation_reward = -6
        for k, z in enumerate(("zone1", "zone2")):
            assert cfg[z]["distribution"] == "normal"
            c.zone_mu[k], c.zone_sigma[k] = cfg[z]["loc"], cfg[z]["scale"]
        c.spawn_robots = _locations(c.N, cfg["end_goal_width"], height, cfg["LEFT"] + cfg["end_goal_width"],
                                    cfg["start_dist"])
    elif scenario == "ArcticTransport":
        c.N = cfg["n_agents"]
        assert c.N == 4
        c.step_dist, c.fast_step, c.slow_step = cfg["normal_step"], cfg["fast_step"], cfg["slow_step"]
        c.not_reached_penalty, c.dist_multiplier = cfg["not_reached_penalty"], cfg["dist_multiplier"]
        c.violation_reward = -30
    elif scenario == "Simple":
        c.N = cfg["n_agents"]
        c.step_dist, c.reward_scaler = cfg["step_dist"], cfg["reward_scaler"]
        c.violation_reward = -5
        c.spawn_robots = _locations(c.N, cfg["ROBOT_INIT_RIGHT_THRESH"] - cfg["LEFT"], height,
                                    cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["start_dist"])
        c.spawn_other = _locations(1, cfg["RIGHT"] - cfg["PREY_INIT_LEFT_THRESH"], height,
                                   cfg["ROBOT_INIT_RIGHT_THRESH"], cfg["step_dist"], spawn_left=False)
    return c


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class COracle(object):
    """Batched driver around orc_step_batch / orc_reset_batch."""

    def __init__(self, scenario, cfg):
        self.scenario = scenario
        self.cfg = dict(cfg)
        self.c = make_config(scenario, cfg)
        L = lib()
        self.N = self.c.N
        self.P = self.c.num_prey
        self.nf, self.ni, self.D = L.orc_nf(self.c), L.orc_ni(self.c), L.orc_obs_dim(self.c)
        self.n_actions = N_ACTIONS[scenario]

    # ---------------------------------------------------------------- dict <-> flat
    def pack(self, st):
        N, P, s = self.N, self.P, self.scenario
        poses = np.asarray(st["poses"], dtype=np.float64)
        batched = poses.ndim == 3
        B = poses.shape[0] if batched else 1

        def get(k, shape, dt):
            return np.asarray(st[k], dtype=dt).reshape((B,) + shape)
        sf = np.zeros((B, self.nf))
        si = np.zeros((B, self.ni), dtype=np.int32)
        sf[:, :3 * N] = get("poses", (3 * N,), np.float64)
        sf[:, 3 * N:6 * N] = get("prev_pose", (3 * N,), np.float64)
        si[:, 0] = get("episode_steps", (), np.int32)
        si[:, 1] = get("prev_valid", (), np.int32)
        if "episode_count" in st:
            si[:, 2] = get("episode_count", (), np.int32)
        if s == "PredatorCapturePrey":
            sf[:, 6 * N:] = get("prey_loc", (2 * P,), np.float64)
            si[:, 3:3 + P] = get("prey_sensed", (P,), np.int32)
            si[:, 3 + P:] = get("prey_captured", (P,), np.int32)
        elif s == "Warehouse":
            si[:, 3:] = get("loaded", (N,), np.int32)
        elif s == "MaterialTransport":
            si[:, 3:3 + N] = get("load", (N,), np.int32)
            si[:, 3 + N:5 + N] = get("zone_load", (2,), np.int32)
            si[:, 5 + N:] = get("messages", (4,), np.int32)
        elif s == "ArcticTransport":
            si[:, 3:99] = get("grid", (96,), np.int32)
            si[:, 99] = get("goal_col", (), np.int32)
            si[:, 100:100 + N] = get("pixel_type", (N,), np.int32)
            si[:, 100 + N:] = get("reached_goal", (N,), np.int32)
        elif s == "Simple":
            sf[:, 6 * N:] = get("goal", (2,), np.float64)
        return sf, si, batched

    def unpack(self, sf, si, batched=True):
        N, P, s = self.N, self.P, self.scenario
        B = sf.shape[0]
        st = {"poses": sf[:, :3 * N].reshape(B, 3, N).copy(),
              "prev_pose": sf[:, 3 * N:6 * N].reshape(B, 3, N).copy(),
              "episode_steps": si[:, 0].copy(), "prev_valid": si[:, 1].copy(),
              "episode_count": si[:, 2].copy()}
        if s == "PredatorCapturePrey":
            st["prey_loc"] = sf[:, 6 * N:].reshape(B, P, 2).copy()
            st["prey_sensed"] = si[:, 3:3 + P].astype(np.uint8)
            st["prey_captured"] = si[:, 3 + P:].astype(np.uint8)
        elif s == "Warehouse":
            st["loaded"] = si[:, 3:].astype(np.uint8)
        elif s == "MaterialTransport":
            st["load"] = si[:, 3:3 + N].copy()
            st["zone_load"] = si[:, 3 + N:5 + N].copy()
            st["messages"] = si[:, 5 + N:].copy()
        elif s == "ArcticTransport":
            st["grid"] = si[:, 3:99].reshape(B, 8, 12).astype(np.uint8)
            st["goal_col"] = si[:, 99].copy()
            st["pixel_type"] = si[:, 100:100 + N].copy()
            st["reached_goal"] = si[:, 100 + N:].astype(np.uint8)
        elif s == "Simple":
            st["goal"] = sf[:, 6 * N:].copy()
        if not batched:
            st = {k: v[0] for k, v in st.items()}
        return st

    # ---------------------------------------------------------------- flat drivers
    def step_flat(self, sf, si, actions, auto_reset=False, seed=0, env_id0=0, threads=1):
        """In place on (sf, si).  Returns obs[B,N,D], reward[B,N], dist[B,N], out_i[B,6]."""
        B = sf.shape[0]
        actions = np.ascontiguousarray(actions, dtype=np.int32).reshape(B, self.N)
        obs = np.empty((B, self.N, self.D))
        rew = np.empty((B, self.N))
        dist = np.empty((B, self.N))
        out_i = np.empty((B, 6), dtype=np.int32)
        L = lib()
        dp, ip = C.c_double, C.c_int32

        def run(lo, hi):
            if hi > lo:
                L.orc_step_batch(self.c, hi - lo, _ptr(sf[lo:], dp), _ptr(si[lo:], ip), _ptr(actions[lo:], ip),
                                 _ptr(obs[lo:], dp), _ptr(rew[lo:], dp), _ptr(dist[lo:], dp), _ptr(out_i[lo:], ip),
                                 int(auto_reset), int(seed), int(env_id0 + lo))
        _threaded(run, B, threads)
        return obs, rew, dist, out_i

    def reset_flat(self, B, seed=0, env_id0=0, sf=None, si=None, threads=1):
        if sf is None:
            sf = np.zeros((B, self.nf))
            si = np.zeros((B, self.ni), dtype=np.int32)
        L = lib()

        def run(lo, hi):
            if hi > lo:
                L.orc_reset_batch(self.c, hi - lo, int(seed), int(env_id0 + lo),
                                  _ptr(sf[lo:], C.c_double), _ptr(si[lo:], C.c_int32))
        _threaded(run, B, threads)
        return sf, si

    def reset_envs(self, sf, si, envs, seed=0, env_id0=0):
        """Re-sample the listed envs in place (what auto_reset does inside step_flat, as a separate call)."""
        L = lib()
        for b in envs:
            b = int(b)
            L.orc_reset_batch(self.c, 1, int(seed), int(env_id0 + b), _ptr(sf[b:], C.c_double), _ptr(si[b:], C.c_int32))

    # ---------------------------------------------------------------- dict drivers (tests)
    def step(self, st, actions):
        """Same contract as RefEnv.step_from: (outputs, post-state), single env or batched."""
        sf, si, batched = self.pack(st)
        obs, rew, dist, out_i = self.step_flat(sf, si, actions)
        out = {"obs": obs, "reward": rew, "dist": dist, "message": out_i[:, 0].copy(),
               "done": out_i[:, 1].astype(np.bool_), "remaining": out_i[:, 2].copy(),
               "qp_evals": out_i[:, 3].copy(), "qp_iters": out_i[:, 4].copy(), "qp_max_iters": out_i[:, 5].copy()}
        if not batched:
            out = {k: v[0] for k, v in out.items()}
            out["done"] = np.full(self.N, out["done"], dtype=np.bool_)
        return out, self.unpack(sf, si, batched)

    def reset(self, B, seed=0, env_id0=0):
        sf, si = self.reset_flat(B, seed, env_id0)
        return self.unpack(sf, si, True)


def _threaded(run, B, threads):
    threads = max(1, min(int(threads), B))
    if threads == 1:
        run(0, B)
        return
    # fine-grained chunks so that envs that early-exit do not unbalance the threads
    chunk = max(1, min(max(256, B // (threads * 8)), -(-B // threads)))
    spans = [(lo, min(B, lo + chunk)) for lo in range(0, B, chunk)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda s: run(*s), spans))


def barrier_qp(dxi, xi, barrier_default=False):
    dxi = np.ascontiguousarray(dxi, dtype=np.float64)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    N = dxi.shape[1]
    u = np.empty((2, N))
    it = lib().orc_barrier_qp(N, int(barrier_default), _ptr(dxi, C.c_double), _ptr(xi, C.c_double), _ptr(u, C.c_double))
    return u, it


def controller(poses, goals, barrier_default=False):
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    goals = np.ascontiguousarray(goals[:2], dtype=np.float64)
    N = poses.shape[1]
    dxu = np.empty((2, N))
    it = lib().orc_controller(N, int(barrier_default), _ptr(poses, C.c_double), _ptr(goals, C.c_double), _ptr(dxu, C.c_double))
    return dxu, it


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(v) for v in o]
