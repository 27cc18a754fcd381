#!/bin/bash
mkdir -p gpurun_out/r2j; O=gpurun_out/r2j
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="--steps 50 --warmup 5 --blocks 3 --no-cpu-baseline --no-extract"
for g in 0 2 4 8 16 32 64; do FW_GROUP_TILES=$g timeout 300 python bench.py --workload c3 $B > $O/b_c3_g$g.json 2>/dev/null; show c3_g$g $O/b_c3_g$g.json; done
for v in s4 s6; do for g in 0 8 32; do FW_GROUP_TILES=$g FW_B200_LIB=$PWD/build_variants/libfw_$v.so timeout 300 python bench.py --workload c3 $B > $O/b_c3_${v}_g$g.json 2>/dev/null; show c3_${v}_g$g $O/b_c3_${v}_g$g.json; done; done
for g in 0 8 32; do FW_GROUP_TILES=$g timeout 300 python bench.py --workload c3r $B > $O/b_c3r_g$g.json 2>/dev/null; show c3r_g$g $O/b_c3r_g$g.json; done
for g in 0 8; do FW_GROUP_TILES=$g timeout 300 python bench.py --workload c2 $B > $O/b_c2_g$g.json 2>/dev/null; show c2_g$g $O/b_c2_g$g.json; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
