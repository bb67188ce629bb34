"""Scratch: host-path step time vs chunk count, and the box's raw pinned PCIe bandwidth."""
import sys, os, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
if len(sys.argv) > 1:
    from marbler_b200 import config
    from marbler_b200.vec_env import VecEnv
    cfg = config.load_yaml(config.default_config_path("PredatorCapturePrey"))
    B = 65536
    env = VecEnv("PredatorCapturePrey", cfg, num_envs=B, device="cuda:0", seed=0, auto_reset=True)
    env.reset()
    h = env.host_buffers()
    a = np.random.RandomState(0).randint(0, 5, size=(B, 4)).astype(np.int32)
    h["actions"].numpy()[...] = a
    for _ in range(5): env.step_host(h["actions"])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): env.step_host(h["actions"])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print("chunks %s: %.3f ms/step  %.3e env-steps/s  D2H %.1f GB/s" % (sys.argv[1], dt * 1e3, B / dt, env.d2h_bytes_per_step / dt / 1e9))
else:
    x = torch.empty(18 * 1024 * 1024, dtype=torch.uint8, device="cuda:0"); y = torch.empty_like(x, device="cpu").pin_memory()
    for n in (1, 8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            for c in range(n):
                k = x.numel() // n
                y[c * k:(c + 1) * k].copy_(x[c * k:(c + 1) * k], non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print("raw D2H 18 MiB in %d copies: %.3f ms  %.1f GB/s" % (n, dt * 1e3, x.numel() / dt / 1e9))
    for c in (1, 2, 4, 8, 16):
        subprocess.run([sys.executable, __file__, str(c)], env=dict(os.environ, MRB_HOST_CHUNKS=str(c)))
