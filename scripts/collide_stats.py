"""Probe: statistics of the collision sweep (library built with -DFW_COLLIDE_STATS, FW_B200_LIB=...)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bevy_firework_b200._native import Engine
from bevy_firework_b200 import workloads as W

wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
eng = Engine(device=0, seed=W.SEED)
sc = bench.Scene(eng, wl, 0)
for _ in range(sc.fill_frames + 20):
    sc.step()
f = eng._L.fw_debug_collide_stats
f.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
out = (C.c_ulonglong * 16)()
import torch
torch.cuda.synchronize()
f(out, 1)
N = 10
for _ in range(N):
    sc.step()
torch.cuda.synchronize()
f(out, 0)
names = ["rays", "rays_on_grid", "warp_casts", "rays_nonempty_cell", "exact_tests", "warp_exact_rounds", "warp_enum_rounds",
         "rays_with_exact", "rays_hit", "warp_casts_with_exact"]
live = eng.total_live()
print("live", live, "frames", N)
for n, v in zip(names, out):
    print(f"{n:24s} {v / N:12.1f} per frame   {v / N / max(live, 1):.4f} per particle")
