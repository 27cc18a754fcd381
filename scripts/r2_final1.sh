#!/bin/bash
# the driver's own commands on one GPU: smoke, default bench line (with parity check + CPU baseline), reference arm
mkdir -p gpurun_out/r2z; O=gpurun_out/r2z
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 2500 $O/bench_default.json; tail -4 $O/bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 3 ) > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 900 $O/bench_reference.json; tail -3 $O/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 40 --csv --log-file $O/launches_c3.csv python bench.py --workload c3 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
for w in c1 c2 c4 c5 c3r c3g; do python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; tail -c 300 $O/bench_$w.json | head -c 300; echo; done
