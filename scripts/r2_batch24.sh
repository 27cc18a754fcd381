#!/bin/bash
mkdir -p gpurun_out/r2b4; O=gpurun_out/r2b4
for r in 30 500 2000 4000 8000 15625 31250 62500; do timeout 120 python scripts/small_scene_probe.py 64 $r 2000 1; done 2>&1 | tee $O/scale64.txt
