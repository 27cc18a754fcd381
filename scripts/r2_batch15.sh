#!/bin/bash
mkdir -p gpurun_out/r2q; O=gpurun_out/r2q
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q/bench_default.json").read().strip().splitlines()[0])
print({k:d[k] for k in ("value","ms_per_step","parity_checked")}, d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"])
print(d["e2e_extract"])
PY
tail -3 $O/bench_default.err
