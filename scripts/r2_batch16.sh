#!/bin/bash
mkdir -p gpurun_out/r2r; O=gpurun_out/r2r
timeout 300 python bench.py --workload c5d --no-cpu-baseline --no-extract --steps 100 > $O/bench_c5d.json 2> $O/bench_c5d.err; tail -c 1500 $O/bench_c5d.json; tail -3 $O/bench_c5d.err
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 140 -c 1 -o $O/c5d python bench.py --workload c5d --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 30 --csv --log-file $O/launches_c5d.csv python bench.py --workload c5d --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ls $O
