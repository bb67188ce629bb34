"""Scratch: the policy kernel on the three model shapes the reference ships (tests/golden/policy), 262,144 agents each,
and the HBM time of its compulsory traffic (hidden state in + out, observations in, actions out)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_util as gu
from marbler_b200.policy import Policy

B, N = 65536, 4
for name, D in (("PredatorCapturePrey_vdn", 16), ("ArcticTransport_qmix_ns", 30), ("MaterialTransport_mappo_ns", 9)):
    z = np.load(os.path.join(gu.GOLDEN, "policy", name + ".npz"))
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    for accurate in (False, True):
        pol = Policy(sd, N, D, device="cuda:0", accurate=accurate)
        obs = torch.rand((B, N, D), device="cuda:0") * 2 - 1
        hid = pol.init_hidden(B)
        act = torch.zeros((B, N), dtype=torch.int32, device="cuda:0")
        for _ in range(5): pol.act(obs, hid, actions=act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): pol.act(obs, hid, actions=act)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        byts = B * N * (2 * 4 * pol.hidden_dim + 4 * D + 4)
        print("%-28s hidden %3d %-6s actions %2d %s: %.4f ms per %d agents (compulsory HBM traffic %.0f MB = %.3f ms at 6.5 TB/s)" % (
            name, pol.hidden_dim, "GRU" if pol.use_rnn else "Linear", pol.n_actions, "float32 mode" if accurate else "tensor cores ", ms, B * N, byts / 1e6, byts / 6.5e9))
