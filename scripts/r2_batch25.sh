#!/bin/bash
mkdir -p gpurun_out/r2b5; O=gpurun_out/r2b5
( FW_FUZZ_SEEDS=5-16 timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python bench.py --workload c5 --no-cpu-baseline --no-extract > $O/bench_c5.json 2> $O/bench_c5.err; python - <<PY
import json
d=json.loads(open("$O/bench_c5.json").read().strip().splitlines()[-1])
print("c5", d["value"], d["ms_per_step"], d["kernel_ms"])
PY
