#!/bin/bash
mkdir -p gpurun_out/r2last; O=gpurun_out/r2last
( timeout 200 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
FW_FUZZ_SEEDS=63-70 timeout 100 python -m pytest tests/test_gpu_fuzz_nested.py -m gpu -q > $O/fuzz_nested_63_70.log 2>&1; tail -3 $O/fuzz_nested_63_70.log
timeout 60 python bench.py --no-cpu-baseline --no-extract > $O/bench_c3.json 2> $O/bench_c3.err; python - <<PY
import json
d=json.loads(open("$O/bench_c3.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
PY
