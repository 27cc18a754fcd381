#!/bin/bash
mkdir -p gpurun_out/r2k; O=gpurun_out/r2k
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
timeout 600 python -m pytest tests -m gpu -x -q -k "collision or textures or mixed or c5 or compaction or lifetime or layout" 2>&1 | tail -3
for w in c5 c3r; do timeout 300 python bench.py --workload $w $B > $O/b_$w.json 2>/dev/null; show $w $O/b_$w.json; done
FW_B200_LIB=$PWD/build_variants/libfw_col2.so timeout 300 python bench.py --workload c5 $B > $O/b_c5_col2.json 2>/dev/null; show c5_col2 $O/b_c5_col2.json
for g in 3 5 7 9 13 27 0; do FW_GROUP_TILES=$g timeout 300 python bench.py --workload c3 $B > $O/b_c3_g$g.json 2>/dev/null; show c3_g$g $O/b_c3_g$g.json; done
ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 140 -c 1 -o $O/c5_coop python bench.py --workload c5 --steps 5 --warmup 3 --blocks 1 --no-cpu-baseline --no-extract --no-graphs > /dev/null 2>&1
