#!/bin/bash
mkdir -p gpurun_out/r2b7; O=gpurun_out/r2b7
for v in col2 col4; do
FW_B200_LIB=build_variants/libfw_$v.so python bench.py --workload c5 --no-cpu-baseline --no-extract > $O/bench_c5_$v.json 2> $O/bench_c5_$v.err; python - <<PY
import json
d=json.loads(open("$O/bench_c5_$v.json").read().strip().splitlines()[-1])
print("$v c5", d["value"], d["ms_per_step"], d["kernel_ms"], "parity", d.get("parity_checked"))
PY
done
