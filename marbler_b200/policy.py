"""On-device policies and rollouts.

The reference evaluates a trained EPyMARL agent on the host, one env, one step at a time
(utilities/misc.py:134-221 run_env): append the one-hot agent id when the model config says `obs_agent_id`
(misc.py:161-162), `q, hs = model(obs, hs)`, `actions = argmax(q)`, `env.step(actions)`; `hs` is re-zeroed at
the start of every episode (misc.py:156).  `Policy` holds the same network (utilities/rnn_agent.py RNNAgent,
utilities/rnn_ns_agent.py RNNNSAgent; weights in their state_dict layout) behind `mrb_policy_act`, a fused CUDA
kernel over all agents of all envs, and `Rollout` chains it with `mrb_step` on one stream - optionally captured
in a CUDA graph - so a whole evaluation runs without a host round trip per step.
PyTorch is used for device memory and streams only; there is no CPU path.
"""
import ctypes as C
import json

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _is_ns(state_dict):
    return any(k.startswith("agents.") for k in state_dict)


def flatten_state_dict(state_dict, n_agents):
    """state_dict of RNNAgent / RNNNSAgent -> (flat float32 array in mrb_policy_create's order, facts)."""
    sd = {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32)
          for k, v in state_dict.items()}
    ns = _is_ns(sd)
    prefixes = ["agents.%d." % i for i in range(n_agents)] if ns else [""]
    use_rnn = (prefixes[0] + "rnn.weight_ih") in sd
    names = ["fc1.weight", "fc1.bias"] + \
        (["rnn.weight_ih", "rnn.weight_hh", "rnn.bias_ih", "rnn.bias_hh"] if use_rnn else ["rnn.weight", "rnn.bias"]) + \
        ["fc2.weight", "fc2.bias"]
    for p in prefixes:
        for n in names:
            if p + n not in sd:
                raise KeyError("state_dict has no %r (n_agents = %d)" % (p + n, n_agents))
    flat = np.concatenate([sd[p + n].reshape(-1) for p in prefixes for n in names]).astype(np.float32)
    hidden, input_dim = sd[prefixes[0] + "fc1.weight"].shape
    return flat, dict(non_shared=ns, use_rnn=use_rnn, hidden_dim=int(hidden), input_dim=int(input_dim),
                      n_actions=int(sd[prefixes[0] + "fc2.weight"].shape[0]))


class Policy(object):
    """Greedy (argmax-q) policy on the GPU.  obs_dim / n_agents describe the env's obs buffer [B, N, D]."""

    def __init__(self, state_dict, n_agents, obs_dim, obs_agent_id=None, device=None, accurate=False):
        """accurate=True: float32 arithmetic like the reference's own forward pass (csrc/policy_f32.cuh) - the greedy
        actions of a checkpoint are then the reference's, at ~10x the time of the FP16 tensor-core kernels."""
        if not torch.cuda.is_available():
            raise RuntimeError("marbler_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        flat, f = flatten_state_dict(state_dict, n_agents)
        if obs_agent_id is None:                         # misc.py:161: the model config decides; infer when not given
            obs_agent_id = f["input_dim"] == obs_dim + n_agents
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.n_agents, self.obs_dim, self.hidden_dim, self.n_actions = n_agents, obs_dim, f["hidden_dim"], f["n_actions"]
        self.obs_agent_id, self.use_rnn, self.non_shared = bool(obs_agent_id), f["use_rnn"], f["non_shared"]
        d = _lib.PolicyDesc(C.sizeof(_lib.PolicyDesc), obs_dim, f["input_dim"], f["hidden_dim"], f["n_actions"], n_agents,
                            int(self.obs_agent_id), int(self.use_rnn), int(self.non_shared), int(bool(accurate)))
        self.accurate = bool(accurate)
        self.handle = C.c_void_p()
        _lib.check_policy(self.lib.mrb_policy_create(C.byref(d), index, flat.ctypes.data_as(C.c_void_p), flat.size,
                                                     C.byref(self.handle)))

    @classmethod
    def from_files(cls, weights_path, model_config_path, n_agents, obs_dim, device=None, accurate=False):
        """The reference's own artefacts: a `.th` state_dict and the sacred model json next to it
        (scenarios/<S>/models/, loaded by utilities/misc.py:66-95 load_env_and_model)."""
        sd = torch.load(weights_path, map_location="cpu")
        with open(model_config_path) as fh:
            mc = json.load(fh)
        return cls(sd, n_agents, obs_dim, obs_agent_id=bool(mc.get("obs_agent_id", False)), device=device, accurate=accurate)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mrb_policy_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def init_hidden(self, num_envs):
        """rnn_agent.py:17-19 init_hidden: zeros, one row per agent."""
        return torch.zeros((num_envs, self.n_agents, self.hidden_dim), dtype=torch.float32, device=self.device)

    def act(self, obs, hidden, actions=None, q=None, fresh=None):
        """obs f32 [B,N,D], hidden f32 [B,N,H] (updated in place) -> actions i32 [B,N] (greedy).  `fresh` u8 [B]:
        envs that start an episode (zero hidden state and observation).  Stream-ordered, no synchronisation."""
        B = obs.shape[0]
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and tuple(obs.shape) == (B, self.n_agents, self.obs_dim)
        assert hidden.dtype == torch.float32 and hidden.is_contiguous() and tuple(hidden.shape) == (B, self.n_agents, self.hidden_dim)
        if actions is None:
            actions = torch.empty((B, self.n_agents), dtype=torch.int32, device=self.device)
        if fresh is not None:
            assert fresh.dtype == torch.uint8 and fresh.is_contiguous() and fresh.shape == (B,)
        with torch.cuda.device(self.device):
            _lib.check_policy(self.lib.mrb_policy_act(
                self.handle, B, _ptr(obs), _ptr(hidden), _ptr(actions), _ptr(q), _ptr(fresh),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), self.handle)
        return actions


class Rollout(object):
    """run_env's loop for a whole batch, on the device: act -> step -> act -> ...  (misc.py:155-187).

    The env auto-resets finished envs inside the step kernel; their `done` flag doubles as the next step's
    `fresh` mask (zero hidden state, zero observation - what run_env feeds after env.reset())."""

    def __init__(self, vec_env, policy, use_graph=True, steps_per_graph=8):
        assert vec_env.N == policy.n_agents and vec_env.D == policy.obs_dim and vec_env.device == policy.device
        assert vec_env.n_actions == policy.n_actions, "policy and scenario disagree on the number of actions"
        self.env, self.policy = vec_env, policy
        self.hidden = policy.init_hidden(vec_env.B)
        self.actions = torch.zeros((vec_env.B, vec_env.N), dtype=torch.int32, device=vec_env.device)
        self.use_graph, self.steps_per_graph = use_graph, int(steps_per_graph)
        self._graph = None

    def reset(self):
        self.env.reset()
        self.env.done.fill_(1)          # every env starts an episode: first act() sees fresh = 1 everywhere
        self.hidden.zero_()

    def _one(self):
        self.policy.act(self.env.obs, self.hidden, actions=self.actions, fresh=self.env.done)
        self.env.step(self.actions)

    def run(self, steps):
        """Advance every env by `steps` env steps (2 kernels per step, no host synchronisation)."""
        k = self.steps_per_graph
        if self.use_graph:
            if self._graph is None and steps > k:
                self._one()                                   # warm-up outside capture
                steps -= 1
                torch.cuda.synchronize(self.env.device)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):           # capture records k steps; it does not run them
                    for _ in range(k):
                        self._one()
            while self._graph is not None and steps >= k:
                self._graph.replay()
                steps -= k
        for _ in range(steps):
            self._one()
