"""Scratch: per-tile timeline of the persistent tcgen05 policy kernel (library built with -DMRB_TC2_TRACE)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_util as gu
from marbler_b200.policy import Policy
from marbler_b200 import _lib
z = np.load(os.path.join(gu.GOLDEN, "policy", "PredatorCapturePrey_vdn.npz"))
sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
B = 65536
pol = Policy(sd, 4, 16, device="cuda:0")
obs = torch.randn(B, 4, 16, device="cuda:0"); hid = pol.init_hidden(B)
for _ in range(3): pol.act(obs, hid)
torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_ulonglong * 256)()
L.mrb_debug_tc2_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
print("rc", L.mrb_debug_tc2_trace(buf))
t = np.array(buf, dtype=np.int64).reshape(16, 16)
t0 = t[0, 10]
names = {10: "stage start", 11: "stage loaded", 12: "act_free seen", 13: "act_full arrive", 5: "mma act_full", 6: "mma free0", 7: "mma p0 issued",
         8: "mma free1", 9: "mma p1 issued", 1: "epi full0", 2: "epi done0", 3: "epi full1", 4: "epi done1"}
order = [10, 11, 12, 13, 5, 6, 7, 8, 9, 1, 2, 3, 4]
print("tile " + " ".join("%15s" % names[k] for k in order))
for n in range(14):
    print("%4d " % n + " ".join("%15.2f" % ((t[n, k] - t0) / 1e3) for k in order))

L.mrb_debug_tc2_trace2.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
L.mrb_debug_tc2_trace2(buf)
t2 = np.array(buf, dtype=np.int64).reshape(16, 16)
print("epilogue, first group of a tile (us since 'old -> tile'): barrier seen | c0: tmem ready, gates done, fc2 done | c1: ... | stores issued")
for n in range(4, 12):
    print("%4d " % n + " ".join("%8.2f" % ((t2[n, k] - t2[n, 0]) / 1e3) for k in range(1, 9)))
