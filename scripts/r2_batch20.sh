#!/bin/bash
mkdir -p gpurun_out/r2v; O=gpurun_out/r2v
for n in 64 512; do
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 24 --csv --log-file $O/launches_small$n.csv python scripts/small_scene_probe.py $n 30 60 0 > /dev/null 2>&1
done
python - <<'PY'
import csv
for f in ("small512","small64"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r2v/launches_{f}.csv")) if len(r)>5 and r[0].isdigit()]
    for r in rows[:8]: print(f, r[4][:70], r[-1])
PY
# kernel-by-kernel timing by the library's own events (profile mode) on c1-like scenes
python - <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from bevy_firework_b200 import _abi, workloads as W
from bevy_firework_b200._native import Engine, frame_input
DT = float(np.float32(1.0) / np.float32(60.0))
for n in (1, 64, 512):
    eng = Engine(device=0, seed=1, profile=True, graphs=False)
    sp = W.stress_spawner(rate=30.0)
    ins = []
    for i, p in enumerate(W.grid_positions(n)):
        ps, nt, es, ne = sp.pods(); eng.spawner_reset(1 + i, ps, nt, es, ne, True); ins.append(frame_input(1 + i, p))
    arr = (_abi.fw_spawner_frame_input * n)(*ins)
    for _ in range(100): eng._L.fw_frame(eng._ctx, DT, arr, n)
    eng.sync(); eng.reset_profile() if hasattr(eng, "reset_profile") else None
    for _ in range(500): eng._L.fw_frame(eng._ctx, DT, arr, n)
    eng.sync()
    print(n, eng.profile() if hasattr(eng, "profile") else None)
PY
