#!/usr/bin/env python
"""Many streams, few particles: n spawners at `rate` particles/s each; frames back to back.
usage: small_scene_probe.py N RATE FRAMES [graphs(0/1)]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevy_firework_b200 import _abi
from bevy_firework_b200._native import Engine, frame_input
from bevy_firework_b200 import workloads as W

n, rate, frames = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
graphs = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
DT = float(np.float32(1.0) / np.float32(60.0))
eng = Engine(device=0, seed=1, graphs=graphs)
sp = W.stress_spawner(rate=rate)
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
for _ in range(frames):
    eng._L.fw_frame(eng._ctx, DT, arr, n)
eng.sync()
t1 = time.perf_counter()
print(f"spawners {n} rate {rate} graphs {graphs}: {1e6*(t1-t0)/frames:.2f} us/frame, live {eng.total_live()}")
