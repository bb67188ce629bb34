"""Scratch: lock-step one config, print the first env whose obs differ."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import c_oracle, golden_util as gu
from marbler_b200.vec_env import VecEnv
np.set_printoptions(precision=5, suppress=True, linewidth=200)
scn = "PredatorCapturePrey"
cfg = dict(gu.Golden(scn + "_rollout").cfg); cfg.update(dict(predator=2, capture=1, num_neighbors=1, num_prey=9))
B = 1024
env = VecEnv(scn, cfg, num_envs=B, device="cuda:0", seed=5, auto_reset=True)
orc = c_oracle.COracle(scn, cfg)
env.reset(); sf, si = orc.reset_flat(B, seed=5, threads=8)
rng = np.random.RandomState(0)
for t in range(3):
    a = rng.randint(0, orc.n_actions, size=(B, orc.N)).astype(np.int32)
    env.step(torch.as_tensor(a, device=env.device))
    obs, rew, dist, out_i = orc.step_flat(sf, si, a, auto_reset=True, seed=5, threads=8)
    go = env.obs.cpu().numpy()
    eo = np.abs(go - obs).reshape(B, -1).max(axis=1)
    bad = np.where(eo > 1e-4)[0]
    print("t", t, "bad", len(bad), bad[:10])
    for b in bad[:3]:
        st = env.get_state(); ost = orc.unpack(sf, si)
        print("env", b, "done", out_i[b, 1], "msg", out_i[b, 0])
        print("gpu obs\n", go[b], "\norc obs\n", obs[b])
        print("gpu poses\n", st["poses"][b], "\norc poses\n", ost["poses"][b])
        print("prey", ost["prey_loc"][b].reshape(-1, 2).T, "captured", ost["prey_captured"][b], st["prey_captured"][b])
    if len(bad): break
