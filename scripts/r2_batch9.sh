#!/bin/bash
mkdir -p gpurun_out/r2i; O=gpurun_out/r2i
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
for w in c3 c3r c4 c2; do timeout 300 python bench.py --workload $w $B > $O/b_$w.json 2> $O/b_$w.err; show $w $O/b_$w.json; done
for v in nograd cs1 cs2 cs3; do FW_B200_LIB=$PWD/build_variants/libfw_$v.so timeout 300 python bench.py --workload c3 $B > $O/b_c3_$v.json 2>/dev/null; show c3_$v $O/b_c3_$v.json; done
for v in cs1 cs2; do FW_B200_LIB=$PWD/build_variants/libfw_$v.so timeout 300 python bench.py --workload c3r $B > $O/b_c3r_$v.json 2>/dev/null; show c3r_$v $O/b_c3r_$v.json; done
timeout 600 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -2
