"""EPyMARL `_GymmaWrapper` semantics of marbler_b200.epymarl.GymmaVecEnv: team reward =
sum over agents, terminated = all(done) or the time limit, zero-padded observations, state = concat(obs)."""
import numpy as np
import pytest
import torch

from marbler_b200 import spaces
from marbler_b200.epymarl import GymmaVecEnv


class _FakeScenario(object):
    def __init__(self, outer):
        self.outer = outer
        self.reset_masks = []

    def reset(self, mask=None, seed=None):
        self.reset_masks.append(mask.clone())
        self.outer.obs_buf[mask] = 0            # like reset_kernel: the re-sampled envs' rows of the obs buffer are zeroed


class _FakeWrapper(object):
    """Stands in for marbler_b200.Wrapper: 3 agents, obs widths 4/6/4, action counts 5/5/3, scripted dones."""

    def __init__(self, num_envs, done_at):
        self.num_envs, self.n_agents, self.done_at, self.t = num_envs, 3, done_at, 0
        self.action_space = spaces.Tuple((spaces.Discrete(5), spaces.Discrete(5), spaces.Discrete(3)))
        self.observation_space = spaces.Tuple(tuple(spaces.Box(-1, 1, (w,), np.float32) for w in (4, 6, 4)))
        self.env = _FakeScenario(self)
        self.obs_buf = torch.zeros((num_envs, 3, 4)) if num_envs > 1 else None      # ONE buffer, rewritten in place

    def reset(self):
        self.t = 0
        if self.num_envs == 1:
            return [[0] * 4] * 3
        self.obs_buf.zero_()
        return self.obs_buf

    def step(self, actions):
        self.t += 1
        if self.num_envs == 1:
            obs = tuple(np.full(4, self.t + i, dtype=np.float32) for i in range(3))
            return obs, [1.0, 2.0, 3.5], [self.t == self.done_at] * 3, {}
        B = self.num_envs
        obs = self.obs_buf.fill_(float(self.t))
        rew = torch.tensor([[1.0, 2.0, 3.5]]).repeat(B, 1)
        done = (torch.arange(B) == self.done_at).unsqueeze(1).expand(B, 3) & (self.t == 2)
        return obs, rew, done, {"message": torch.zeros(B, dtype=torch.uint8), "remaining": torch.zeros(B, dtype=torch.int32)}


def test_single_env_matches_gymma_semantics():
    env = GymmaVecEnv("robotarium_gym:PredatorCapturePrey-v0", time_limit=5, env=_FakeWrapper(1, done_at=3))
    obs, state = env.reset()
    assert len(obs) == 3 and all(o.shape == (6,) and not o.any() for o in obs) and state.shape == (18,)
    assert env.get_env_info() == {"state_shape": 18, "obs_shape": 6, "n_actions": 5, "n_agents": 3, "episode_limit": 5}
    assert env.get_avail_actions() == [[1] * 5, [1] * 5, [1, 1, 1, 0, 0]]
    r, term, info = env.step([0, 1, 2])
    assert isinstance(r, float) and r == 6.5 and term is False and info == {}
    assert np.array_equal(env.get_obs_agent(1), np.array([2, 2, 2, 2, 0, 0], dtype=np.float32))
    assert np.array_equal(env.get_state(), np.concatenate(env.get_obs()))
    env.step([0, 0, 0])
    assert env.step([0, 0, 0])[1] is True                     # the env's own done
    env.reset()
    for _ in range(4):
        assert env.step([0, 0, 0])[1] is False or env._wrapper.t == 3
    assert env.step([0, 0, 0])[1] is True                     # the time limit (5 steps)


def test_batched_time_limit_and_reset_mask():
    B = 4
    env = GymmaVecEnv("robotarium_gym:Warehouse-v0", time_limit=3, num_envs=B, env=_FakeWrapper(B, done_at=1))
    obs, state = env.reset()
    assert obs.shape == (B, 3, 6) and state.shape == (B, 18) and env.get_avail_actions().shape == (B, 3, 5)
    a = torch.zeros((B, 3), dtype=torch.int32)
    r, term, info = env.step(a)
    assert torch.equal(r, torch.full((B,), 6.5)) and not term.any()
    r, term, info = env.step(a)                               # env 1 finishes on its own at t = 2
    assert term.tolist() == [False, True, False, False] and not info["TimeLimit.truncated"].any()
    assert env._elapsed.tolist() == [2, 0, 2, 2]
    assert info["fresh"].tolist() == [False, True, False, False]
    assert env.get_obs()[1, 0, 0] == 2.0 and env.get_obs(episode_start=True)[1, 0, 0] == 0.0
    r, term, info = env.step(a)                               # the others hit the limit at t = 3
    assert term.tolist() == [True, False, True, True]
    assert info["TimeLimit.truncated"].tolist() == [True, False, True, True]
    assert env._env.reset_masks[-1].tolist() == [True, False, True, True]
    assert env._elapsed.tolist() == [0, 1, 0, 0]
    assert torch.equal(env.get_state(), env.get_obs().reshape(B, -1))
    assert float(env.get_obs()[..., 4:].abs().max()) == 0.0   # zero padding of the short observations
    # episode boundaries: the terminal observation (t = 3) of the truncated envs survives the masked reset that zeroed
    # their rows of the env's own buffer, and the own-done env (reset inside the step at t = 2) reports the same way
    assert float(env._wrapper.obs_buf[0].abs().max()) == 0.0
    assert env.get_obs()[:, 0, 0].tolist() == [3.0, 3.0, 3.0, 3.0]
    assert info["fresh"].tolist() == [True, False, True, True] and torch.equal(env.fresh_mask(), info["fresh"])
    assert env.get_obs(episode_start=True)[:, 0, 0].tolist() == [0.0, 3.0, 0.0, 0.0]
    assert env.get_state(episode_start=True)[:, 0].tolist() == [0.0, 3.0, 0.0, 0.0]


@pytest.mark.gpu
def test_adapter_on_the_cuda_env():
    B = 256
    env = GymmaVecEnv("robotarium_gym:PredatorCapturePrey-v0", time_limit=6, num_envs=B, seed=4)
    obs, state = env.reset()
    assert obs.shape == (B, 4, 16) and state.shape == (B, 64) and obs.is_cuda
    gen = torch.Generator(device="cuda").manual_seed(0)
    ended = torch.zeros(B, dtype=torch.bool, device="cuda")
    for t in range(6):
        a = torch.randint(0, 5, (B, 4), generator=gen, device="cuda", dtype=torch.int32)
        r, term, info = env.step(a)
        assert torch.allclose(r, env._env.vec.reward.sum(dim=1))
        ended |= term
    assert bool(term.any()) and bool(ended.all())             # everybody is cut at the limit at the latest
    # truncated envs were re-sampled by mrb_reset (their rows of the device obs buffer are zero now); the adapter still
    # reports their terminal observations, and zeros as the first observation of the next episode
    trunc = info["TimeLimit.truncated"]
    assert bool(trunc.any()) and float(env._env.vec.obs[trunc].abs().max()) == 0.0
    assert bool((env.get_obs()[trunc].abs().amax(dim=(1, 2)) > 0).all())
    assert float(env.get_obs(episode_start=True)[term].abs().max()) == 0.0
    single = GymmaVecEnv("robotarium_gym:MaterialTransport-v0", time_limit=4, seed=1)
    obs, state = single.reset()
    assert len(obs) == 4 and state.shape == (36,)
    r, term, _ = single.step([0, 5, 10, 19])
    assert isinstance(r, float) and isinstance(term, bool)
