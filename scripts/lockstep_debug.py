"""Scratch: run GPU env and C oracle in lockstep, on first divergence dump the offending env."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import c_oracle, golden_util as gu
from marbler_b200.vec_env import VecEnv, barrier_qp

scn = sys.argv[1] if len(sys.argv) > 1 else "PredatorCapturePrey"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
T = int(sys.argv[3]) if len(sys.argv) > 3 else 60
g = gu.Golden(scn + "_rollout")
env = VecEnv(scn, g.cfg, num_envs=B, device="cuda:0", seed=5, auto_reset=True)
orc = c_oracle.COracle(scn, g.cfg)
env.reset(); sf, si = orc.reset_flat(B, seed=5, threads=8)
rng = np.random.RandomState(0)
nbad = 0
for t in range(T):
    a = rng.randint(0, orc.n_actions, size=(B, orc.N)).astype(np.int32)
    pre = orc.unpack(sf.copy(), si.copy())
    gpre = env.get_state()
    env.step(torch.as_tensor(a, device=env.device))
    obs, rew, dist, out_i = orc.step_flat(sf, si, a, auto_reset=True, seed=5, threads=8)
    eo = np.abs(env.obs.cpu().numpy() - obs).reshape(B, -1).max(axis=1)
    bad = np.where((eo > 1e-5) | (env.message.cpu().numpy() != out_i[:, 0]) | (env.done.cpu().numpy() != out_i[:, 1]))[0]
    if len(bad):
        b = bad[0]
        print("t=%d divergent envs %s  obs err %.3e  msg gpu %d orc %d  oracle qp evals %d iters %d" % (
            t, bad[:8], eo[b], env.message[b].item(), out_i[b, 0], out_i[b, 3], out_i[b, 4]))
        print("pre-state equal:", {k: bool(np.allclose(np.asarray(gpre[k][b], dtype=float), np.asarray(pre[k][b], dtype=float), atol=1e-12)) for k in pre if k in gpre})
        one = {k: v[b:b + 1] for k, v in pre.items()}
        np.savez(os.path.join(ROOT, "gpurun_out", "diverge_%s.npz" % scn), actions=a[b], **one)
        # replay this env alone on both sides
        e1 = VecEnv(scn, g.cfg, num_envs=1, device="cuda:0", seed=5)
        e1.set_state(one); e1.step(torch.as_tensor(a[b:b + 1], device="cuda:0")); torch.cuda.synchronize()
        o1, s1 = orc.step(one, a[b:b + 1])
        print("replay: gpu poses\n", e1.get_state()["poses"][0], "\noracle poses\n", s1["poses"][0])
        print("gpu stats", e1.read_stats())
        # first controller evaluation inputs
        p = one["poses"][0]; N = orc.N
        xi = p[:2] + 0.05 * np.stack([np.cos(p[2]), np.sin(p[2])])
        print("min pair dist", (np.hypot(xi[0][:, None] - xi[0][None], xi[1][:, None] - xi[1][None]) + 9 * np.eye(N)).min())
        nbad += 1
        if nbad >= 2: break
        # resync the oracle to the GPU state so that we can look for further, independent divergences
        st = env.get_state(); sf, si, _ = orc.pack(st)
print("done", t)
