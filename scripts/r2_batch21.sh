#!/bin/bash
mkdir -p gpurun_out/r2w; O=gpurun_out/r2w
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for n in 1 64 512 2048; do timeout 120 python scripts/small_scene_probe.py $n 30 3000 1; done 2>&1 | tee $O/small.txt
timeout 300 python bench.py --no-cpu-baseline --no-extract > $O/bench_c3.json 2> $O/bench_c3.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2w/bench_c3.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d.get("parity_checked"))
PY
