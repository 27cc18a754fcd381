import time, sys, numpy as np
sys.path.insert(0, '.')
import bench
from bevy_firework_b200._native import Engine
from bevy_firework_b200 import workloads as W
eng = Engine(device=0, seed=W.SEED)
sc = bench.Scene(eng, sys.argv[1] if len(sys.argv) > 1 else 'c3', 0)
for _ in range(sc.fill_frames + 20): sc.step()
eng.sync()
keys = [k for k, *_ in sc.spawners]
tf = tc = ta = ts = 0.0
N = 200
for _ in range(N):
    t0 = time.perf_counter(); sc.step(); t1 = time.perf_counter()
    eng.sync(); t2 = time.perf_counter()
    eng.counts_all(); t3 = time.perf_counter()
    eng.read_aabb(keys[0]); t4 = time.perf_counter()
    tf += t1 - t0; ts += t2 - t1; tc += t3 - t2; ta += t4 - t3
print(f"per step us: fw_frame {tf/N*1e6:.1f}  sync-wait {ts/N*1e6:.1f}  counts_all {tc/N*1e6:.1f}  read_aabb {ta/N*1e6:.1f}")
