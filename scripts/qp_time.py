"""Scratch: time the stand-alone barrier-QP kernel (mrb_barrier_qp) on problems taken from a running env, to compare
the solver's speed inside and outside the fused step kernel.  python scripts/qp_time.py [scenario] [B]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marbler_b200 import _lib, config
from marbler_b200.vec_env import VecEnv

scn = sys.argv[1] if len(sys.argv) > 1 else "PredatorCapturePrey"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
cfg = config.load_yaml(config.default_config_path(scn))
env = VecEnv(scn, cfg, num_envs=B, device="cuda:0", seed=0, auto_reset=True)
env.reset()
gen = torch.Generator(device="cuda:0").manual_seed(0)
for _ in range(20):
    env.step(torch.randint(0, env.n_actions, (B, env.N), generator=gen, device="cuda:0", dtype=torch.int32))
N = env.N
p = env.state_f64[:3 * N].clone()                       # rows x[N] | y[N] | th[N], env fastest
x, y, th = p[:N], p[N:2 * N], p[2 * N:]
xi = torch.cat([x + 0.05 * torch.cos(th), y + 0.05 * torch.sin(th)]).contiguous()         # [2N][B]
step = float(cfg.get("step_dist", 0.2))
a = torch.randint(0, 5, (N, B), generator=gen, device="cuda:0")
gx = torch.where(a == 0, x - step, torch.where(a == 1, x + step, x))
gy = torch.where(a == 2, y - step, torch.where(a == 3, y + step, y))
d = torch.cat([gx, gy]) - xi
nrm = torch.sqrt(d[:N] ** 2 + d[N:] ** 2).clamp_min(1e-12)
sc = torch.where(nrm > 0.15, 0.15 / nrm, torch.ones_like(nrm))
dxi = (d * torch.cat([sc, sc])).contiguous()
u = torch.empty_like(dxi)
it = torch.zeros(B, dtype=torch.int32, device="cuda:0")
lib = _lib.load()
vp = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    _lib.check(lib.mrb_barrier_qp(0, N, 0, B, vp(dxi), vp(xi), vp(u), vp(it), st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    lib.mrb_barrier_qp(0, N, 0, B, vp(dxi), vp(xi), vp(u), vp(it), st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
itf = it.float()
wmax = itf.view(-1, 32).max(dim=1).values.mean().item()
print("%s N=%d: QP kernel alone %.4f ms for %d problems; iterations mean %.2f, warp-max mean %.2f -> %.2f ns per warp-iteration-slot "
      "(time x resident warps / (warps x warp-max))" % (scn, N, ms, B, itf.mean().item(), wmax, ms * 1e6 / (B / 32 * wmax)))
