"""Minimal stand-in for the old-API `gym` package (gym 0.21 era), ORACLE ONLY.

Test infrastructure: lets /root/reference/robotarium_gym import unmodified in a
container where gym is not installed.  Restated from the public gym API
(SURVEY.md App. A.10): Env.reset()->obs, Env.step()->(obs, reward, done, info),
spaces.{Discrete,Box,Tuple}, envs.registration.register / gym.make.
"""
from . import spaces  # noqa: F401
from .envs import registration as _registration
from .envs.registration import register, make  # noqa: F401


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        pass

    def close(self):
        pass

    def seed(self, seed=None):
        return [seed]
