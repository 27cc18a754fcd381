#!/bin/bash
mkdir -p gpurun_out/r2d; O=gpurun_out/r2d
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
tail -8 $O/pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], r["kernel"], "B/p", r["algorithmic_bytes_per_particle"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], {k:round(v,4) for k,v in d["kernel_ms"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
B="--steps 50 --warmup 5 --blocks 5 --no-cpu-baseline --no-extract"
for w in c3 c3g c3r c4 c5 c2 c1; do timeout 300 python bench.py --workload $w $B > $O/bench_$w.json 2> $O/bench_$w.err; show $w $O/bench_$w.json; done
for v in s6 s4 s5np s6np; do FW_B200_LIB=$PWD/build_variants/libfw_$v.so timeout 300 python bench.py --workload c3 $B > $O/bench_c3_$v.json 2> $O/bench_c3_$v.err; show c3_$v $O/bench_c3_$v.json; done
FW_B200_LIB=$PWD/build_variants/libfw_r4.so timeout 300 python bench.py --workload c3g $B > $O/bench_c3g_r4.json 2> $O/bench_c3g_r4.err; show c3g_r4 $O/bench_c3g_r4.json
FW_B200_LIB=$PWD/build_variants/libfw_s6.so timeout 300 python bench.py --workload c4 $B > $O/bench_c4_s6.json 2> $O/bench_c4_s6.err; show c4_s6 $O/bench_c4_s6.json
