#!/bin/bash
mkdir -p gpurun_out/r2o; O=gpurun_out/r2o
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="--steps 100 --warmup 5 --blocks 3 --no-cpu-baseline --no-extract"
timeout 300 python bench.py --workload c3 $B > $O/b_c3.json 2>/dev/null; show c3 $O/b_c3.json
for v in u2 u2s4 u2s3; do FW_B200_LIB=$PWD/build_variants/libfw_$v.so timeout 300 python bench.py --workload c3 $B > $O/b_c3_$v.json 2>/dev/null; show c3_$v $O/b_c3_$v.json; done
FW_B200_LIB=$PWD/build_variants/libfw_u2.so timeout 300 python bench.py --workload c3r $B > $O/b_c3r_u2.json 2>/dev/null; show c3r_u2 $O/b_c3r_u2.json
FW_B200_LIB=$PWD/build_variants/libfw_u2s4.so timeout 300 python bench.py --workload c4 $B > $O/b_c4_u2s4.json 2>/dev/null; show c4_u2s4 $O/b_c4_u2s4.json
python scripts/host_cost_probe.py
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -3
FW_B200_LIB=$PWD/build_variants/libfw_u2.so timeout 900 python -m pytest tests -m gpu -x -q -k "layout or fullsize or mixed" 2>&1 | tail -3
