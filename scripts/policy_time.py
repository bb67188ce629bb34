"""Scratch: time the policy kernel and the graph-replayed rollout."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_util as gu
from marbler_b200.policy import Policy, Rollout
from marbler_b200.vec_env import VecEnv
z = np.load(os.path.join(gu.GOLDEN, "policy", "PredatorCapturePrey_vdn.npz"))
sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
cfg = dict(gu.Golden("PredatorCapturePrey_rollout").cfg)
B = 65536
env = VecEnv("PredatorCapturePrey", cfg, num_envs=B, device="cuda:0", seed=0, auto_reset=True)
pol = Policy(sd, 4, 16, device="cuda:0", accurate="accurate" in sys.argv[1:])
ro = Rollout(env, pol, use_graph=True, steps_per_graph=16)
ro.reset()
ro.run(40)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for _ in range(50): pol.act(env.obs, ro.hidden, actions=ro.actions, fresh=env.done)
ev[1].record()
for _ in range(50): env.step(ro.actions)
ev[2].record()
ro.run(16 * 20)
ev[3].record()
torch.cuda.synchronize()
print("policy kernel %.4f ms  step kernel %.4f ms  rollout %.4f ms/step (graph, %d envs)" % (
    ev[0].elapsed_time(ev[1]) / 50, ev[1].elapsed_time(ev[2]) / 50, ev[2].elapsed_time(ev[3]) / 320, B))
print(env.read_stats())
