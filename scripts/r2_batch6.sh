#!/bin/bash
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
FW_GROUP_TILES=0 ncu --set full --clock-control none --import-source on -k regex:update_stream_kernel -s 80 -c 1 -o $O/c3_stream python bench.py --workload c3 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
FW_GROUP_TILES=0 ncu --set full --clock-control none --import-source on -k regex:update_stream_kernel -s 140 -c 1 -o $O/c5_stream python bench.py --workload c5 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ls -la $O
