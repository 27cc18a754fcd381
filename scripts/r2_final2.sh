#!/bin/bash
# final evidence on the last code: GPU suite, smoke, default bench line, reference arm, every workload
mkdir -p gpurun_out/r2z2; O=gpurun_out/r2z2
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 1200 $O/bench_default.json; tail -4 $O/bench_default.err
( time python bench.py --impl reference --steps 20 --warmup 3 ) > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json
for w in c1 c2 c4 c5 c5d c3r c3g; do python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; python - <<PY
import json
d=json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["kernel_ms"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", d.get("parity_checked"))
PY
done
