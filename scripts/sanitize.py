"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python scripts/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from marbler_b200 import config
from marbler_b200.vec_env import VecEnv

def run(scn, B, over=None, steps=3, host=False):
    cfg = config.load_yaml(config.default_config_path(scn))
    cfg.update(over or {})
    env = VecEnv(scn, cfg, num_envs=B, device="cuda:0", seed=3, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for _ in range(steps):
        a = torch.randint(0, env.n_actions, (B, env.N), generator=g, device="cuda:0", dtype=torch.int32)
        if host:
            env.step_host(a.cpu())
        else:
            env.step(a)
    torch.cuda.synchronize()
    print("ok", scn, B, over, "host" if host else "device")

for scn in ("PredatorCapturePrey", "Warehouse", "MaterialTransport", "ArcticTransport", "Simple"):
    run(scn, 200)
run("PredatorCapturePrey", 40, dict(predator=10, capture=10, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))
run("PredatorCapturePrey", 70, dict(predator=4, capture=4, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))
run("PredatorCapturePrey", 33000, host=True, steps=2)
run("Warehouse", 300, host=True, steps=2)
if os.environ.get("SAN_POLICY", "1") == "1":
    import golden_util as gu
    from marbler_b200.policy import Policy, Rollout
    z = np.load(os.path.join(gu.GOLDEN, "policy", "PredatorCapturePrey_vdn.npz"))
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    cfg = dict(gu.Golden("PredatorCapturePrey_rollout").cfg)
    env = VecEnv("PredatorCapturePrey", cfg, num_envs=300, device="cuda:0", seed=0, auto_reset=True)
    pol = Policy(sd, 4, 16, device="cuda:0")
    ro = Rollout(env, pol, use_graph=False)
    ro.reset(); ro.run(3)
    torch.cuda.synchronize()
    print("ok policy", os.environ.get("MRB_POLICY_TC", "0"))
