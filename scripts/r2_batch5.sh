#!/bin/bash
mkdir -p gpurun_out/r2e; O=gpurun_out/r2e
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="--steps 50 --warmup 5 --blocks 3 --no-cpu-baseline --no-extract"
for g in 0 2 4 8 16 32; do
  for w in c3 c3g c3r; do FW_GROUP_TILES=$g timeout 300 python bench.py --workload $w $B > $O/b_${w}_g$g.json 2> $O/b_${w}_g$g.err; show ${w}_g$g $O/b_${w}_g$g.json; done
done
for g in 4 8 16; do
  FW_GROUP_TILES=$g FW_B200_LIB=$PWD/build_variants/libfw_r4.so timeout 300 python bench.py --workload c3g $B > $O/b_c3g_r4_g$g.json 2>/dev/null; show c3g_r4_g$g $O/b_c3g_r4_g$g.json
  FW_GROUP_TILES=$g FW_B200_LIB=$PWD/build_variants/libfw_s5np.so timeout 300 python bench.py --workload c3 $B > $O/b_c3_np_g$g.json 2>/dev/null; show c3_np_g$g $O/b_c3_np_g$g.json
done
for w in c5 c4 c2 c1; do timeout 300 python bench.py --workload $w $B > $O/b_$w.json 2>/dev/null; show $w $O/b_$w.json; done
timeout 600 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -3
