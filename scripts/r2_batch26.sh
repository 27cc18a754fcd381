#!/bin/bash
mkdir -p gpurun_out/r2b6; O=gpurun_out/r2b6
( timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for w in c3 c3g c5 c3r c2; do python bench.py --workload $w --no-cpu-baseline --no-extract > $O/bench_$w.json 2> $O/bench_$w.err; python - <<PY
import json
d=json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["kernel_ms"], "e2e", d["e2e"]["value"], "parity", d.get("parity_checked"))
PY
done
