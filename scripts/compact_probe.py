import sys, numpy as np
sys.path.insert(0, '.')
import bench
from bevy_firework_b200._native import Engine
from bevy_firework_b200 import workloads as W
wl = sys.argv[1] if len(sys.argv) > 1 else 'c3r'
eng = Engine(device=0, seed=W.SEED)
sc = bench.Scene(eng, wl, 0)
for _ in range(sc.fill_frames + 20): sc.step()
eng.sync()
for mode in ("back-to-back", "sync after every frame"):
    eng.profile_reset(); eng.set_profiling(True)
    for _ in range(100):
        sc.step()
        if mode != "back-to-back": eng.sync()
    eng.sync()
    p, n = eng.profile_sum()
    eng.set_profiling(False)
    print(wl, mode, "update_ms %.4f spawn_ms %.4f frame %.4f" % (p.update_ms / 100, p.spawn_ms / 100, p.total_ms / 100))
