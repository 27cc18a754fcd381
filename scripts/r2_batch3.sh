#!/bin/bash
mkdir -p gpurun_out/r2c; O=gpurun_out/r2c
( time timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > $O/pytest_gpu.log 2>&1
tail -15 $O/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 80 -c 1 -o $O/c3_static python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"update_kernel|count_kernel|scan_kernel" -s 330 -c 3 -o $O/c3r_static python bench.py --workload c3r --steps 5 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
ls -la $O
