#!/bin/bash
mkdir -p gpurun_out/r2l; O=gpurun_out/r2l
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
B="--steps 100 --warmup 5 --blocks 5 --no-cpu-baseline --no-extract"
for w in c3 c3r c3g c4 c5 c2 c1; do timeout 300 python bench.py --workload $w $B > $O/b_$w.json 2> $O/b_$w.err; show $w $O/b_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 40 --csv --log-file $O/launches_c3r.csv python bench.py --workload c3r --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file $O/launches_c3.csv python bench.py --workload c3 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:update_static_kernel -s 80 -c 1 -o $O/c3_static_v3 python bench.py --workload c3 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"update_static_kernel|count_kernel" -s 220 -c 2 -o $O/c3r_static_v3 python bench.py --workload c3r --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
