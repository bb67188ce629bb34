"""compute-sanitizer exercise of the warp kernels only (tensor-core 20-robot solver, generic team sizes, QP alone):
   compute-sanitizer --tool racecheck python scripts/sanitize_warp.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from marbler_b200 import config
from marbler_b200.vec_env import VecEnv, barrier_qp

def run(B, over, steps=2):
    cfg = config.load_yaml(config.default_config_path("PredatorCapturePrey"))
    cfg.update(over)
    env = VecEnv("PredatorCapturePrey", cfg, num_envs=B, device="cuda:0", seed=3, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for _ in range(steps):
        env.step(torch.randint(0, env.n_actions, (B, env.N), generator=g, device="cuda:0", dtype=torch.int32))
    torch.cuda.synchronize()
    print("ok", B, over)

run(24, dict(predator=10, capture=10, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))          # 20 robots, folded
run(16, dict(predator=4, capture=5, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))            # 9 on 12, padded
run(8, dict(predator=12, capture=11, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3), steps=1)  # 23 on 24
run(5, dict(predator=15, capture=14, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3), steps=1)  # 29 on 32
os.environ["MRB_WARP_GENERIC"] = "1"                                                          # run-time team size kernels
run(16, dict(predator=4, capture=5, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3))
run(8, dict(predator=12, capture=11, ROBOT_INIT_RIGHT_THRESH=0.1, num_neighbors=3), steps=1)
del os.environ["MRB_WARP_GENERIC"]
g = torch.Generator(device="cuda:0").manual_seed(2)
xi = torch.rand((16, 2, 20), generator=g, device="cuda:0", dtype=torch.float64) * 2 - 1
dxi = torch.rand((16, 2, 20), generator=g, device="cuda:0", dtype=torch.float64) * 0.4 - 0.2
u, it = barrier_qp(dxi, xi)
u2, it2 = barrier_qp(dxi[:, :, :18].contiguous(), xi[:, :, :18].contiguous())              # 18 on 20, padded
torch.cuda.synchronize()
print("ok qp", it.cpu().numpy().tolist(), it2.cpu().numpy().tolist())
