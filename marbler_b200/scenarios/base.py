"""Scenario contract (reference: robotarium_gym/scenarios/base.py:1-30) and the batched
implementation shared by the five scenarios.  The arithmetic of step / reset is in the CUDA
kernels (csrc/); these classes keep the reference's method names, argument meaning and return
types so existing callers (wrapper.Wrapper, EPyMARL's gymma wrapper) work unchanged.

num_envs == 1 -> the reference's own Python structures (lists / tuples / numpy arrays / info dict);
num_envs  > 1 -> tensors: obs [B,N,D] f32, reward [B,N] f32, done [B,N] bool, info dict of tensors.
"""
import numpy as np
import torch

from .. import spaces
from ..config import objectview, obs_space_dim
from ..reference_rng import sample_reset
from ..vec_env import VecEnv

MESSAGES = ("", "collision", "boundary", "collision_boundary")      # roboEnv.py:85-90


class BaseEnv(object):
    """Template of a Robotarium environment; must also expose num_robots, agent_poses, visualizer."""

    def get_action_space(self):
        raise NotImplementedError()

    def get_observation_space(self):
        raise NotImplementedError()

    def step(self, actions_):
        raise NotImplementedError()

    def reset(self):
        raise NotImplementedError()

    def render(self, mode="human"):
        pass

    def _generate_step_goal_positions(self, actions):
        raise NotImplementedError()


class _NoVisualizer(object):
    """Rendering is out of scope (show_figure_frequency = -1 on the batched path)."""
    show_figure = False


class BatchedScenario(BaseEnv):
    scenario = None
    obs_low, obs_high = -1.5, 3.0

    def __init__(self, args, num_envs=1, device=None, seed=None, env_id0=0, auto_reset=None,
                 track_dist=True, collect_stats=True, reset_rng=None):
        if isinstance(args, dict):
            args = objectview(dict(args))
        self.args = args
        cfg = args.__dict__
        if cfg.get("show_figure_frequency", -1) != -1 or cfg.get("save_gif", False):
            # visualisation / gif capture is not part of the batched step path
            cfg = dict(cfg, show_figure_frequency=-1, save_gif=False)
        self.num_envs = int(num_envs)
        # reset sampler: "reference" = numpy's global legacy RNG + Python's random, call for call as the
        # reference draws (seed-for-seed episodes, single env only); "philox" = the in-kernel counter-based
        # stream (same distribution; batched, sharding-invariant)
        if reset_rng is None:
            reset_rng = "reference" if self.num_envs == 1 else "philox"
        if reset_rng not in ("reference", "philox"):
            raise ValueError("reset_rng must be 'reference' or 'philox'")
        if reset_rng == "reference" and self.num_envs != 1:
            raise ValueError("reset_rng='reference' reproduces the reference's single global RNG stream: num_envs must be 1")
        self.reset_rng = reset_rng
        self._cfg = dict(cfg)
        if reset_rng == "reference" and seed is None and cfg.get("seed", -1) != -1:
            np.random.seed(cfg["seed"])         # e.g. PredatorCapturePrey.py:27-28 (global side effect, as the reference)
        if seed is None:                        # the reference seeds numpy when seed != -1 (e.g. PredatorCapturePrey.py:27-28)
            seed = cfg.get("seed", -1)
            seed = int(np.random.SeedSequence().entropy & (2 ** 63 - 1)) if seed == -1 else int(seed)
        if auto_reset is None:
            auto_reset = self.num_envs > 1
        self.vec = VecEnv(self.scenario, cfg, num_envs=self.num_envs, device=device, seed=seed, env_id0=env_id0,
                          auto_reset=auto_reset, track_dist=track_dist, collect_stats=collect_stats,
                          f64_outputs=self.num_envs == 1)     # one env: float64 observations / rewards like the reference
        self.num_robots = self.vec.N
        self.num_agent = self.num_robots
        self.visualizer = _NoVisualizer()
        self.env = self                         # the reference keeps its roboEnv here; the sim is fused into step()
        self.episode_steps = 0
        width = obs_space_dim(self.scenario, cfg, self.vec.c)
        self.action_space = spaces.Tuple(tuple(spaces.Discrete(self.vec.n_actions) for _ in range(self.num_robots)))
        self.observation_space = spaces.Tuple(tuple(
            spaces.Box(low=self.obs_low, high=self.obs_high, shape=(width,), dtype=np.float32)
            for _ in range(self.num_robots)))

    # ------------------------------------------------------------------ reference surface
    def get_action_space(self):
        return self.action_space

    def get_observation_space(self):
        return self.observation_space

    @property
    def agent_poses(self):
        p = self.vec.agent_poses
        return p[0].cpu().numpy() if self.num_envs == 1 else p

    def reset(self, mask=None, seed=None):
        if self.reset_rng == "reference":
            st = sample_reset(self.scenario, self._cfg)
            self.vec.set_state({k: np.asarray(v)[None] for k, v in st.items()})
            self.vec.obs.zero_()
            self.vec.obs_f64.zero_()
        else:
            self.vec.reset(mask=mask, seed=seed)
        if self.num_envs == 1:                  # e.g. PredatorCapturePrey.py:136: an all-zero observation
            self.episode_steps = 0
            return [[0] * self.vec.D] * self.num_robots
        return self.vec.obs

    def step(self, actions_):
        if self.num_envs == 1:
            a = np.asarray([int(x) for x in actions_], dtype=np.int32).reshape(1, self.num_robots)
            obs, rew, done, msg = self.vec.step_host(a)
            self.episode_steps += 1
            code, terminated = int(msg[0]), bool(done[0])
            info = self._info_single(code, terminated, int(self.vec.remaining[0].item()))
            if self.vec.dist is not None:
                info["dist_travelled"] = self.vec.dist[0].cpu().numpy().astype(np.float64)
            # float64 like the reference (e.g. PredatorCapturePrey.py:176 returns numpy float64 rows and Python floats)
            o = self.vec.obs_f64[0].cpu().numpy()
            r = self.vec.reward_f64[0].cpu().numpy()
            return [o[i].copy() for i in range(self.num_robots)], [float(v) for v in r], \
                [terminated] * self.num_robots, info
        if isinstance(actions_, torch.Tensor) and actions_.is_cuda:
            obs, rew, done, msg = self.vec.step(actions_)
        else:
            obs, rew, done, msg = self.vec.step_host(actions_)
        info = {"message": msg, "remaining": self.vec.remaining}
        if self.vec.dist is not None:
            info["dist_travelled"] = self.vec.dist
        return obs, rew, done.bool().unsqueeze(1).expand(self.num_envs, self.num_robots), info

    def _info_single(self, code, terminated, remaining):
        info = {}
        if code:
            info["message"] = MESSAGES[code]
        return info

    def get_observations(self, *_):
        """Observations of the most recent step (computed inside the fused step kernel)."""
        if self.num_envs == 1:
            o = self.vec.obs_f64[0].cpu().numpy()
            return [o[i].copy() for i in range(self.num_robots)]
        return self.vec.obs

    def get_rewards(self, *_):
        if self.num_envs == 1:
            return [float(v) for v in self.vec.reward_f64[0].cpu()]
        return self.vec.reward

    get_reward = get_rewards                    # MaterialTransport / ArcticTransport spell it get_reward

    def _step_sizes(self):
        c = self.vec.c
        return torch.full((self.num_envs, self.num_robots), c.step_dist, dtype=torch.float64, device=self.vec.device)

    def _generate_step_goal_positions(self, actions):
        """Goal poses for the current poses and actions (reference: Agent.generate_goal, e.g.
        PredatorCapturePrey/agent.py:48-76).  Host-side helper for callers that want the goals; the
        step kernel computes the same thing internally."""
        c, dev = self.vec.c, self.vec.device
        a = torch.as_tensor(np.asarray(actions) if not isinstance(actions, torch.Tensor) else actions, device=dev)
        a = a.reshape(self.num_envs, self.num_robots).long()
        if self.scenario == "MaterialTransport":
            a = a // 4
        goal = self.vec.agent_poses.clone()
        x, y, s = goal[:, 0], goal[:, 1], self._step_sizes()
        cx, cy = x.clamp(c.left, c.right), y.clamp(c.up, c.down)
        gx = torch.where(a == 0, (x - s).clamp_min(c.left), torch.where(a == 1, (x + s).clamp_max(c.right), cx))
        gy = torch.where(a == 2, (y - s).clamp_min(c.up), torch.where(a == 3, (y + s).clamp_max(c.down), cy))
        goal[:, 0], goal[:, 1] = gx, gy
        return goal[0].cpu().numpy() if self.num_envs == 1 else goal

    # ------------------------------------------------------------------ checkpoint / injection
    def get_state(self):
        return self.vec.get_state()

    def set_state(self, st):
        self.vec.set_state(st)
