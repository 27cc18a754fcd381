#!/bin/bash
mkdir -p gpurun_out/r2x4; O=gpurun_out/r2x4
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for k in "nested_emission_textures or growth or large_one_shot" "layout"; do
timeout 900 compute-sanitizer --tool initcheck --print-limit 4 python -m pytest tests -m gpu -x -q -k "$k" > $O/initcheck.log 2>&1
grep -n "Uninitialized\|at .*fw_\|at .*kernel\|ERROR SUMMARY\|passed\|failed" $O/initcheck.log | head -20
cat $O/initcheck.log | tail -4 >> $O/initcheck_all.log
done
