"""Generate tests/golden/policy/*.npz from the reference's OWN agent classes and shipped checkpoints.

ORACLE / TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference).
    python oracle/gen_policy_golden.py

For a few of the reference's trained models (scenarios/<S>/models/*.th + *.json) the fixture holds the
state_dict, a short sequence of observations, and what utilities/rnn_agent.py RNNAgent / utilities/
rnn_ns_agent.py RNNNSAgent (imported unmodified) return when driven as utilities/misc.py:155-170 run_env drives
them: q values, greedy actions and the final hidden state, in float32 (`q`, `actions`, `h`).  `q_tf32` /
`h_tf32` are the same network evaluated with every matmul operand rounded to FP16 (what the CUDA kernel's
tensor-core MMAs see; same 11-bit significand as TF32, which the first version used - hence the names) - the exact target of the kernel's arithmetic; the float32 numbers are the parity bar.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_harness import _ensure_paths, REFERENCE_ROOT  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "policy")
MODELS = [  # scenario, weights, model json, actor class, n_agents, env obs dim, n_actions
    ("PredatorCapturePrey", "vdn.th", "vdn.json", "RNNAgent", 4, 16, 5),          # shared, H=128, obs_agent_id
    ("ArcticTransport", "qmix_ns.th", "qmix_ns.json", "RNNNSAgent", 4, 30, 5),    # non-shared GRU, H=64
    ("MaterialTransport", "mappo_ns.th", "mappo_ns.json", "RNNNSAgent", 4, 9, 20),  # non-shared, Linear instead of GRU
]


def tf32(x):
    """Operand rounding of the kernel's tensor-core MMAs: FP16, round to nearest even (the first version of
    the kernel used TF32, hence the historical names q_tf32 / h_tf32 of the emulated outputs)."""
    return x.half().float()


def tf32_agent(sd, prefix, use_rnn, x, h):
    """RNNAgent.forward (rnn_agent.py:21-29) with FP16-rounded matmul operands, FP32 accumulation."""
    lin = lambda v, w, b: tf32(v) @ tf32(sd[prefix + w]).t() + sd[prefix + b]
    x = torch.relu(lin(x, "fc1.weight", "fc1.bias"))
    if use_rnn:
        gi, gh = lin(x, "rnn.weight_ih", "rnn.bias_ih"), lin(h, "rnn.weight_hh", "rnn.bias_hh")
        H = h.shape[-1]
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        hn = (1 - z) * n + z * h
    else:
        hn = torch.relu(lin(x, "rnn.weight", "rnn.bias"))
    return lin(hn, "fc2.weight", "fc2.bias"), hn


def main():
    _ensure_paths()
    from robotarium_gym.utilities.rnn_agent import RNNAgent
    from robotarium_gym.utilities.rnn_ns_agent import RNNNSAgent
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    for scn, wfile, jfile, cls, N, D, A in MODELS:
        mdir = os.path.join(REFERENCE_ROOT, "robotarium_gym", "scenarios", scn, "models")
        sd = torch.load(os.path.join(mdir, wfile), map_location="cpu")
        mc = json.load(open(os.path.join(mdir, jfile)))
        args = types.SimpleNamespace(hidden_dim=mc["hidden_dim"], use_rnn=mc["use_rnn"], n_actions=A, n_agents=N)
        input_dim = sd[list(sd.keys())[0]].shape[1]                      # misc.py:83
        model = (RNNAgent if cls == "RNNAgent" else RNNNSAgent)(input_dim, args)
        model.load_state_dict(sd)
        model.eval()
        B, T, H = 48, 4, mc["hidden_dim"]
        lo, hi = (-1.5, 1.5)
        obs = torch.rand(T, B, N, D) * (hi - lo) + lo
        if scn == "MaterialTransport":
            obs[..., 2:] = torch.randint(0, 20, obs[..., 2:].shape).float()
        obs[:, :, :, -1] = torch.round(obs[:, :, :, -1])
        hs = np.zeros((B, N, H), dtype=np.float32)
        ht = torch.zeros(B, N, H)
        qs, acts, qts = [], [], []
        with torch.no_grad():
            for t in range(T):
                q_t, h_t = np.zeros((B, N, A), np.float32), np.zeros((B, N, H), np.float32)
                qt_t = torch.zeros(B, N, A)
                for b in range(B):                                       # one env at a time, exactly like run_env
                    o = obs[t, b].numpy()
                    if mc["obs_agent_id"]:
                        o = np.concatenate([o, np.eye(N)], axis=1)       # misc.py:161-162
                    if cls == "RNNNSAgent":
                        q, h = model(torch.Tensor(o), torch.Tensor(hs[b].T.copy()).t().unsqueeze(0))   # [1, N, H]
                        h = h[0]
                    else:
                        q, h = model(torch.Tensor(o), torch.Tensor(hs[b]))
                    q_t[b], h_t[b] = q.numpy(), h.numpy()
                    for a in range(N):
                        prefix = "agents.%d." % a if cls == "RNNNSAgent" else ""
                        qq, hh = tf32_agent(sd, prefix, mc["use_rnn"], torch.Tensor(o[a:a + 1]), ht[b, a:a + 1].clone())
                        qt_t[b, a], ht[b, a] = qq[0], hh[0]
                hs = h_t
                qs.append(q_t), acts.append(np.argmax(q_t, axis=2)), qts.append(qt_t.numpy().copy())
        blob = {"scenario": np.array(scn), "model": np.array(wfile), "n_agents": np.int32(N), "obs_dim": np.int32(D),
                "obs_agent_id": np.int32(bool(mc["obs_agent_id"])), "obs": obs.numpy(), "q": np.stack(qs),
                "actions": np.stack(acts).astype(np.int32), "h": hs, "q_tf32": np.stack(qts), "h_tf32": ht.numpy()}
        for k, v in sd.items():
            blob["sd." + k] = v.numpy().astype(np.float32)
        path = os.path.join(OUT, "%s_%s.npz" % (scn, wfile[:-3]))
        np.savez_compressed(path, **blob)
        gap = np.sort(np.stack(qs), axis=-1)
        print("%-36s in %d H %d  |q| max %.2f  |q - q_tf32| max %.2e  min top-2 gap %.2e  agree(tf32) %.4f  %6.1f KB" % (
            os.path.basename(path), input_dim, H, np.abs(np.stack(qs)).max(), np.abs(np.stack(qs) - np.stack(qts)).max(),
            (gap[..., -1] - gap[..., -2]).min(), (np.stack(qts).argmax(-1) == np.stack(acts)).mean(), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
