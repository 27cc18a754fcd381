#!/usr/bin/env python
"""Host cost of fw_frame as a function of the number of spawners: scenes with almost no device work
(rate 30 particles/s per spawner), frames submitted back to back, wall clock per frame."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevy_firework_b200 import _abi
from bevy_firework_b200._native import Engine, frame_input
from bevy_firework_b200 import workloads as W

DT = float(np.float32(1.0) / np.float32(60.0))
for n in (1, 64, 512, 2048):
    eng = Engine(device=0, seed=1)
    sp = W.stress_spawner(rate=30.0)
    inputs = []
    for i, p in enumerate(W.grid_positions(n)):
        ps, nt, es, ne = sp.pods()
        eng.spawner_reset(1 + i, ps, nt, es, ne, True)
        inputs.append(frame_input(1 + i, p))
    arr = (_abi.fw_spawner_frame_input * n)(*inputs)
    for _ in range(100):
        eng._L.fw_frame(eng._ctx, DT, arr, n)
    eng.sync()
    t0 = time.perf_counter()
    K = 2000
    for _ in range(K):
        eng._L.fw_frame(eng._ctx, DT, arr, n)
    t1 = time.perf_counter()
    eng.sync()
    t2 = time.perf_counter()
    # host-only cost: 3 submissions into an idle ring (4 slots: no back-pressure), GPU drained in between
    acc = 0.0
    R = 300
    for _ in range(R):
        eng.sync()
        a = time.perf_counter()
        for _ in range(3):
            eng._L.fw_frame(eng._ctx, DT, arr, n)
        acc += time.perf_counter() - a
    host_only = 1e6 * acc / (3 * R)
    print(f"spawners {n:5d}: host-only {host_only:6.2f} us/frame (3 frames into an idle ring)", flush=True)
    print(f"spawners {n:5d}: submit {1e6*(t1-t0)/K:7.2f} us/frame, incl. drain {1e6*(t2-t0)/K:7.2f} us/frame, live {eng.total_live()}", flush=True)
    eng.close()
