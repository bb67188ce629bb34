"""Gym wrapper with the reference's surface (robotarium_gym/wrapper.py:19-50): Wrapper(env_name,
config_path) -> reset() / step(action_n) / observation_space / action_space / n_agents / env.
Extra keyword arguments select the batch size and device; with the defaults (num_envs=1) the return
types are the reference's own."""
import yaml

from .config import default_config_path, objectview
from .scenarios import PredatorCapturePrey, Warehouse, MaterialTransport, ArcticTransport, simple

try:                                            # pragma: no cover - gym is optional
    from gym import Env
except Exception:
    try:                                        # pragma: no cover
        from gymnasium import Env
    except Exception:
        class Env(object):
            metadata = {}

env_dict = {"PredatorCapturePrey": PredatorCapturePrey,
            "Warehouse": Warehouse,
            "MaterialTransport": MaterialTransport,
            "Simple": simple,
            "ArcticTransport": ArcticTransport}


class Wrapper(Env):
    def __init__(self, env_name, config_path=None, num_envs=1, device=None, config_overrides=None, **kwargs):
        super().__init__()
        with open(config_path or default_config_path(env_name), "r") as f:
            config = yaml.safe_load(f)
        if config_overrides:
            config.update(config_overrides)
        args = objectview(config)
        self.env = env_dict[env_name](args, num_envs=num_envs, device=device, **kwargs)
        self.observation_space = self.get_observation_space()
        self.action_space = self.get_action_space()
        self.n_agents = self.env.num_robots
        self.num_envs = num_envs

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action_n):
        obs_n, reward_n, done_n, info_n = self.env.step(action_n)
        if self.num_envs == 1:
            return tuple(obs_n), reward_n, done_n, info_n
        return obs_n, reward_n, done_n, info_n

    def get_action_space(self):
        return self.env.get_action_space()

    def get_observation_space(self):
        return self.env.get_observation_space()
