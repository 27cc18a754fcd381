#!/bin/bash
mkdir -p gpurun_out/r2b3; O=gpurun_out/r2b3
( timeout 1500 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for w in c1 c2 c3 c5; do python bench.py --workload $w --no-cpu-baseline --no-extract > $O/bench_$w.json 2> $O/bench_$w.err; python - <<PY
import json
d=json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["kernel_ms"], "e2e", d["e2e"]["value"], "parity", d.get("parity_checked"))
PY
done
for n in 1 64 512; do timeout 120 python scripts/small_scene_probe.py $n 30 3000 1; done 2>&1 | tee $O/small.txt
