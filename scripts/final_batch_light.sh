#!/bin/bash
# lighter evidence batch (about 5 GPU-minutes): tests, every bench workload, launch list, one ncu capture, sanitizers on the kernels changed last
mkdir -p gpurun_out/final2; O=gpurun_out/final2
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4) > $O/pytest_gpu.log
python bench.py --steps 200 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
for w in c1 c2 c3r c4 c5; do python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 24 --csv --log-file $O/launches_c3r.csv python bench.py --workload c3r --steps 3 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 110 -c 1 -o $O/c3r python bench.py --workload c3r --steps 5 --warmup 3 --no-cpu-baseline --no-graphs > /dev/null 2>&1
(timeout 400 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "random_lifetime_compaction or destroyed_stream or collision_destroy or collision_scene" 2>&1 | tail -4) > $O/racecheck.log
(timeout 400 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "compaction or destroyed or collision or nested" 2>&1 | tail -4) > $O/memcheck.log
cat $O/pytest_gpu.log; tail -2 $O/racecheck.log; tail -2 $O/memcheck.log
for w in c3 c3r c5; do tail -1 $O/bench_$w.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['kernel_ms']['update'], d['roofline']['frac'], d['value'], d['e2e']['value'])"; done
