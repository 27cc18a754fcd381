#!/bin/bash
mkdir -p gpurun_out/r2y; O=gpurun_out/r2y
ncu --set full --clock-control none --import-source on -k regex:update_static -s 70 -c 1 -o $O/c2 python bench.py --workload c2 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ls -la $O
