#!/bin/bash
# round 2, second GPU batch: the per-stream layout (static streams, constant packs not stored)
mkdir -p gpurun_out/r2b
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2b/pytest_gpu.log 2>&1
tail -25 gpurun_out/r2b/pytest_gpu.log
for w in c3 c3r c4 c5 c2 c1; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2b/bench_$w.json 2> gpurun_out/r2b/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2b/bench_$w.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$w", "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], r["kernel"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], d["kernel_ms"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/r2b/bench_$w.err").read()[-800:])
PY
done
