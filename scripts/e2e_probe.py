#!/usr/bin/env python
"""Where an end-to-end step goes: wall time of the fw_frame call itself (host work up to the last
enqueue), of the wait for the frame's results (fw_counts_all), per workload."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bevy_firework_b200._native import Engine
from bevy_firework_b200 import workloads as W

for wl in (sys.argv[1:] or ["c3", "c2", "c1"]):
    eng = Engine(device=0, seed=W.SEED)
    sc = bench.Scene(eng, wl, 0)
    for _ in range(sc.fill_frames + 20):
        sc.step()
    eng.sync()
    t_frame = t_wait = 0.0
    K = 300
    for _ in range(K):
        t0 = time.perf_counter()
        sc.step()
        t1 = time.perf_counter()
        eng.counts_all()
        t2 = time.perf_counter()
        t_frame += t1 - t0
        t_wait += t2 - t1
    print(f"{wl}: fw_frame call {1e6*t_frame/K:7.1f} us, wait for results {1e6*t_wait/K:7.1f} us, step {1e6*(t_frame+t_wait)/K:7.1f} us", flush=True)
    eng.close()
