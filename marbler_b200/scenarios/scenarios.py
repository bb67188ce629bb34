"""The five scenario classes under the reference's names (robotarium_gym/wrapper.py:12-16)."""
import torch

from .base import BatchedScenario, MESSAGES


class PredatorCapturePrey(BatchedScenario):
    """robotarium_gym/scenarios/PredatorCapturePrey/PredatorCapturePrey.py"""
    scenario = "PredatorCapturePrey"
    obs_low, obs_high = -5, 3                                         # PredatorCapturePrey.py:54

    def __init__(self, args, **kw):
        super().__init__(args, **kw)
        self.num_prey = self.vec.P
        self.num_predators = self.vec.c.num_predators
        self.num_capture = self.num_robots - self.num_predators
        self.agent_obs_dim = 6 if self.vec.c.capability_aware else 4

    def _info_single(self, code, terminated, remaining):              # PredatorCapturePrey.py:155-169
        if code:
            return {"message": MESSAGES[code]}
        return {"remaining": remaining} if terminated else {}


class Warehouse(BatchedScenario):
    """robotarium_gym/scenarios/Warehouse/warehouse.py"""
    scenario = "Warehouse"
    obs_low, obs_high = -1.5, 1.5                                     # warehouse.py:71
    agent_obs_dim = 3


class MaterialTransport(BatchedScenario):
    """robotarium_gym/scenarios/MaterialTransport/MaterialTransport.py"""
    scenario = "MaterialTransport"
    obs_low, obs_high = -1.5, 1.5                                     # MaterialTransport.py:81

    def _step_sizes(self):                                            # MaterialTransport.py:71-74
        c = self.vec.c
        s = torch.tensor([c.fast_step if i < c.n_fast else c.slow_step for i in range(self.num_robots)],
                         dtype=torch.float64, device=self.vec.device)
        return s.unsqueeze(0).expand(self.num_envs, -1)

    def _info_single(self, code, terminated, remaining):              # MaterialTransport.py:135-144
        info = {"message": MESSAGES[code]} if code else {}
        if terminated:
            info["remaining"] = remaining
        return info


class ArcticTransport(BatchedScenario):
    """robotarium_gym/scenarios/ArcticTransport/ArcticTransport.py"""
    scenario = "ArcticTransport"
    obs_low, obs_high = -1.5, 3                                       # ArcticTransport.py:47
    agent_obs_dim = 30

    def _step_sizes(self):                                            # ArcticTransport/agent.py:94-112
        c, dev = self.vec.c, self.vec.device
        pix = self.vec.state_i32[10].long()
        out = torch.full((self.num_envs, 4), c.fast_step, dtype=torch.float64, device=dev)
        table = {2: (c.step_dist, c.fast_step, c.slow_step, c.step_dist),     # ice
                 3: (c.step_dist, c.slow_step, c.fast_step, c.step_dist)}     # water
        for i, t in table.items():
            out[:, i] = torch.tensor(t, dtype=torch.float64, device=dev)[(pix >> (2 * i)) & 3]
        return out


class simple(BatchedScenario):
    """robotarium_gym/scenarios/Simple/simple.py (the reference class is lower-case)."""
    scenario = "Simple"
    obs_low, obs_high = -1.5, 3                                       # simple.py:99

    def _info_single(self, code, terminated, remaining):              # simple.py:176 files the message under 'remaining'
        return {"remaining": MESSAGES[code]} if code else {}
