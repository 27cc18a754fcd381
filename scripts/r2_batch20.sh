#!/bin/bash
mkdir -p gpurun_out/r2v; O=gpurun_out/r2v
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 24 --csv --log-file $O/launches_small512.csv python scripts/small_scene_probe.py 512 30 60 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 24 --csv --log-file $O/launches_small64.csv python scripts/small_scene_probe.py 64 30 60 0 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("small512","small64"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r2v/launches_{f}.csv")) if len(r)>5 and r[0].isdigit()]
    for r in rows[:10]: print(f, r[4][:70], r[-1])
PY
