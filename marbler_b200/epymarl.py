"""EPyMARL-facing adapter, batched  (SURVEY.md section 8f-2).

EPyMARL does not talk to `robotarium_gym.wrapper.Wrapper` directly: its `envs/__init__.py` wraps every gym
env in `_GymmaWrapper(key, time_limit, pretrained_wrapper, **kwargs)`, which (restated from EPyMARL; it is
not vendored in the reference, whose README.md:28 only tells the user which `time_limit` to pass)

  * wraps the env in EPyMARL's own `TimeLimit(env, max_episode_steps=time_limit)`: after `time_limit` steps
    `done` is forced to all-True and `info["TimeLimit.truncated"] = not all(done)`;
  * pads every agent's observation with zeros to the longest observation in `observation_space`;
  * `step(actions) -> (float(sum(reward)), all(done), {})`           (team reward = SUM over agents),
  * `get_obs()`, `get_obs_agent(i)`, `get_obs_size()`, `get_state() = concat(obs)`, `get_state_size()`,
    `get_avail_actions()` (all ones up to the agent's own action count, zero padded), `get_total_actions()`,
    `reset() -> (obs, state)`, `get_env_info()`, `get_stats() = {}`, `close()`, `seed()`.

`GymmaVecEnv` offers the same methods over `num_envs` environments at once: every quantity gains a leading
env axis and stays a CUDA tensor (`reward [B]`, `terminated [B]`, `obs [B, N, D]`, `state [B, N*D]`,
`avail_actions [B, N, A]`), so a learner can consume `[B, ...]` batches without EPyMARL's subprocess/Pipe
parallel runner.  With `num_envs == 1` the return types are EPyMARL's own (float, bool, list of numpy
arrays), i.e. it is a drop-in for `_GymmaWrapper`.  Envs that finish (own `done` or the time limit) are
re-sampled in place before the next step; `episode_limit` truncation resets them through `mrb_reset`'s mask.

Episode boundaries in the batched mode (both ways of finishing behave the same): after `step()`, `get_obs()` /
`get_state()` hold the TERMINAL observation of every env that just finished (what EPyMARL's runners store as
the last transition of the episode) - a private copy, so the masked re-sampling that follows (which zeroes
the reset envs' rows of the device obs buffer, like the reference's `reset()` returning zeros) cannot touch
it.  `info["fresh"]` / `fresh_mask()` flag those envs: they start a new episode at the next step, and
`get_obs(episode_start=True)` / `get_state(episode_start=True)` return the buffer with their rows zeroed,
i.e. exactly what the reference's `reset()` would have handed the learner as the first observation.
"""
import numpy as np
import torch

from .wrapper import Wrapper


def _scenario_of(key):
    name = key.split(":")[-1]
    return name[:-3] if name.endswith("-v0") else name


class GymmaVecEnv(object):
    def __init__(self, key, time_limit, num_envs=1, pretrained_wrapper=None, env=None, **kwargs):
        if pretrained_wrapper:
            raise NotImplementedError("pretrained_wrapper is an lbforaging/rware feature EPyMARL never uses with MARBLER")
        self.episode_limit = int(time_limit)
        self._wrapper = env if env is not None else Wrapper(_scenario_of(key), num_envs=num_envs, **kwargs)
        self._env = self._wrapper.env
        self.num_envs = int(getattr(self._wrapper, "num_envs", num_envs))
        self.n_agents = self._wrapper.n_agents
        spaces_a, spaces_o = self._wrapper.action_space, self._wrapper.observation_space
        self.longest_action_space = max(spaces_a, key=lambda x: x.n)
        self.longest_observation_space = max(spaces_o, key=lambda x: x.shape)
        self._obs_size = int(np.prod(self.longest_observation_space.shape))
        self._n_actions = int(self.longest_action_space.n)
        self._obs = None
        self._elapsed = None
        self._truncated = None
        self._fresh = None

    # ------------------------------------------------------------------ helpers
    def _pad(self, obs):
        """Zero-pad the last axis to the declared (longest) observation width."""
        if self.num_envs == 1:
            return [np.pad(np.asarray(o, dtype=np.float32), (0, self._obs_size - len(o)), "constant", constant_values=0)
                    for o in obs]
        d = obs.shape[-1]
        return obs if d == self._obs_size else torch.nn.functional.pad(obs, (0, self._obs_size - d))

    # ------------------------------------------------------------------ MultiAgentEnv surface
    def step(self, actions):
        """Returns (reward, terminated, info): floats / bools for one env, tensors [B] otherwise."""
        if self.num_envs == 1:
            actions = [int(a) for a in actions]
            obs, reward, done, info = self._wrapper.step(actions)
            self._elapsed += 1
            if self._elapsed >= self.episode_limit:                 # EPyMARL TimeLimit.step
                info["TimeLimit.truncated"] = not all(done)
                done = len(obs) * [True]
            self._obs = self._pad(obs)
            return float(sum(reward)), all(done), {}
        obs, reward, done, info = self._wrapper.step(actions)
        self._elapsed += 1
        terminated = done[:, 0].clone()
        limit = self._elapsed >= self.episode_limit
        self._truncated = limit & ~terminated
        terminated |= limit
        team_reward = reward.sum(dim=1)
        # private copy: `obs` is the env's device buffer, and the masked reset below zeroes the rows of the envs it
        # re-samples - the terminal observations must survive it (own-done envs keep theirs in the buffer too)
        self._obs = self._pad(obs).clone()
        self._fresh = terminated
        # own-done envs were re-sampled inside the step kernel (auto-reset); truncated ones are reset here
        if bool(self._truncated.any()):
            self._env.reset(mask=self._truncated)
        self._elapsed = torch.where(terminated, torch.zeros_like(self._elapsed), self._elapsed)
        return team_reward, terminated, {"TimeLimit.truncated": self._truncated, "message": info["message"],
                                         "remaining": info["remaining"], "fresh": self._fresh}

    def fresh_mask(self):
        """[B] bool: envs whose episode ended at the last step (they begin a new one at the next step)."""
        return self._fresh

    def get_obs(self, episode_start=False):
        """List of per-agent observations (one env) or the [B, N, D] tensor.  After a step the rows of finished
        envs are their terminal observations; episode_start=True returns them zeroed instead (the reference's
        reset() returns an all-zero observation, e.g. PredatorCapturePrey.py:136)."""
        if episode_start and self.num_envs > 1 and self._fresh is not None:
            return torch.where(self._fresh.view(-1, 1, 1), torch.zeros_like(self._obs), self._obs)
        return self._obs

    def get_obs_agent(self, agent_id):
        return self._obs[agent_id] if self.num_envs == 1 else self._obs[:, agent_id]

    def get_obs_size(self):
        return self._obs_size

    def get_state(self, episode_start=False):
        if self.num_envs == 1:
            return np.concatenate(self._obs, axis=0).astype(np.float32)
        return self.get_obs(episode_start).reshape(self.num_envs, self.n_agents * self._obs_size)

    def get_state_size(self):
        return self.n_agents * self._obs_size

    def get_avail_actions(self):
        if self.num_envs == 1:
            return [self.get_avail_agent_actions(i) for i in range(self.n_agents)]
        out = torch.zeros((self.num_envs, self.n_agents, self._n_actions), dtype=torch.int64, device=self._obs.device)
        for i, sp in enumerate(self._wrapper.action_space):
            out[:, i, :sp.n] = 1
        return out

    def get_avail_agent_actions(self, agent_id):
        valid = self._wrapper.action_space[agent_id].n * [1]
        return valid + [0] * (self._n_actions - len(valid))

    def get_total_actions(self):
        return self._n_actions

    def reset(self):
        obs = self._wrapper.reset()
        if self.num_envs == 1:
            self._elapsed = 0
            self._obs = self._pad(obs)
        else:
            self._elapsed = torch.zeros((self.num_envs,), dtype=torch.int32, device=obs.device)
            self._obs = self._pad(obs).clone()
            self._fresh = torch.ones((self.num_envs,), dtype=torch.bool, device=obs.device)
        return self.get_obs(), self.get_state()

    def render(self):
        pass

    def close(self):
        pass

    def seed(self):
        return None

    def save_replay(self):
        pass

    def get_stats(self):
        return {}

    def get_env_info(self):
        return {"state_shape": self.get_state_size(), "obs_shape": self.get_obs_size(),
                "n_actions": self.get_total_actions(), "n_agents": self.n_agents,
                "episode_limit": self.episode_limit}
