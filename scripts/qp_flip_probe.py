"""Scratch: compare GPU barrier QP and C oracle on many random problems; dump disagreements."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import c_oracle
from marbler_b200.vec_env import barrier_qp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
rng = np.random.RandomState(0)
# PCP-like: robots in the left part of the arena, min distance 0.21, random nominal velocities
xi = np.stack([rng.uniform(-1.4, -0.2, (B, N)), rng.uniform(-0.9, 0.9, (B, N))], axis=1)
ang = rng.uniform(0, 2 * np.pi, (B, N)); mag = rng.uniform(0, 0.15, (B, N))
dxi = np.stack([mag * np.cos(ang), mag * np.sin(ang)], axis=1)
u, it = barrier_qp(torch.tensor(dxi, device="cuda:0"), torch.tensor(xi, device="cuda:0"))
u = u.cpu().numpy(); it = it.cpu().numpy()
import ctypes as C
uo = np.empty_like(u); ito = np.empty(B, dtype=np.int32)
from concurrent.futures import ThreadPoolExecutor
def run(span):
    for b in range(*span):
        uu, ii = c_oracle.barrier_qp(dxi[b], xi[b]); uo[b] = uu; ito[b] = ii
spans = [(i, min(B, i + 5000)) for i in range(0, B, 5000)]
with ThreadPoolExecutor(16) as ex: list(ex.map(run, spans))
err = np.abs(u - uo).reshape(B, -1).max(axis=1)
flip = it != ito
print("N=%d B=%d  flips %d (%.2e)  max err same-iters %.3e  max err flipped %.3e  mean iters %.2f" % (
    N, B, flip.sum(), flip.mean(), err[~flip].max(), err[flip].max() if flip.any() else 0, it.mean()))
print("err quantiles same-iters:", np.quantile(err[~flip], [0.5, 0.99, 0.9999, 1.0]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
idx = np.where(flip | (err > 1e-7))[0][:200]
np.savez(os.path.join(ROOT, "gpurun_out", "qp_flips_N%d.npz" % N), dxi=dxi[idx], xi=xi[idx], u=u[idx], uo=uo[idx], it=it[idx], ito=ito[idx])
