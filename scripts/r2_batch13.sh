#!/bin/bash
mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
timeout 600 python -m pytest tests -m gpu -x -q -k "compaction or lifetime or mixed or layout or nested or destroyed" 2>&1 | tail -3
B="--steps 100 --warmup 5 --blocks 5 --no-cpu-baseline --no-extract"
timeout 300 python bench.py --workload c3r $B > $O/b_c3r.json 2> $O/b_c3r.err; show c3r $O/b_c3r.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 40 --csv --log-file $O/launches_c3r.csv python bench.py --workload c3r --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
grep -E "count_kernel|scan_kernel|update_static" $O/launches_c3r.csv | awk -F'","' '{print $5, $(NF)}' | head -8
