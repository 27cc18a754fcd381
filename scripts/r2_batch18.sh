#!/bin/bash
mkdir -p gpurun_out/r2t; O=gpurun_out/r2t
for T in 0 256 100000000; do
  FW_INLINE_TILES=$T timeout 200 python bench.py --workload c1 --no-cpu-baseline --no-extract --steps 2000 > $O/c1_T$T.json 2> $O/c1_T$T.err
  python - <<PY
import json
d=json.loads(open("$O/c1_T$T.json").read().strip().splitlines()[-1])
print("c1 T=$T ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d.get("parity_checked"))
PY
done
for T in 0 100000000; do
  FW_INLINE_TILES=$T timeout 200 python bench.py --workload c2 --no-cpu-baseline --no-extract --steps 500 > $O/c2_T$T.json 2> $O/c2_T$T.err
  python - <<PY
import json
d=json.loads(open("$O/c2_T$T.json").read().strip().splitlines()[-1])
print("c2 T=$T ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
FW_INLINE_TILES=100000000 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_inline_all.log 2>&1; tail -3 $O/pytest_inline_all.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; tail -3 $O/pytest_default.log
