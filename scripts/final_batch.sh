#!/bin/bash
# one-GPU evidence batch: tests, every bench workload, reference arm, ncu launch list + captures, sanitizer
mkdir -p gpurun_out; O=gpurun_out/final
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $O/pytest_gpu.log
python bench.py --steps 200 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
for w in c1 c2 c3r c4 c5; do python bench.py --workload $w --steps 200 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err; done
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_c3.json 2> $O/bench_reference_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 110 -c 1 -o $O/c3r python bench.py --workload c3r --steps 5 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 80 -c 1 -o $O/c3 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 140 -c 1 -o $O/c5 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
(timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "compaction or nested or destroyed or collision_destroy or collision_single or (randomized_mixed_scene and 1)" 2>&1 | tail -6) > $O/racecheck.log
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "compaction or nested or destroyed or random or collision or edge" 2>&1 | tail -6) > $O/memcheck.log
cat $O/pytest_gpu.log; tail -2 $O/racecheck.log; tail -2 $O/memcheck.log
