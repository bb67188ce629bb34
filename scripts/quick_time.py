"""Scratch timing of the step kernel (not the bench): python scripts/quick_time.py [scenario] [B] [steps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marbler_b200 import config
from marbler_b200.vec_env import VecEnv

scn = sys.argv[1] if len(sys.argv) > 1 else "PredatorCapturePrey"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
over = {}
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    over[k] = eval(v)
cfg = config.load_yaml(config.default_config_path(scn))
cfg.update(over)
env = VecEnv(scn, cfg, num_envs=B, device="cuda:0", seed=0, auto_reset=True)
env.reset()
gen = torch.Generator(device="cuda:0").manual_seed(0)
acts = [torch.randint(0, env.n_actions, (B, env.N), generator=gen, device="cuda:0", dtype=torch.int32) for _ in range(steps + 5)]
for i in range(5):
    env.step(acts[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    env.step(acts[5 + i])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
st = env.read_stats()
print("%s N=%d B=%d: %.3f ms/step  %.3e env-steps/s  %.3e agent-steps/s  iters/qp %.2f  episodes %d" % (
    scn, env.N, B, ms, B / ms * 1e3, B * env.N / ms * 1e3, st["qp_iterations"] / max(st["qp_solves"], 1), st["episodes"]))
