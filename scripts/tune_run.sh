#!/bin/bash
# run the C3 bench once per tuning variant and print the update-kernel time
for f in build_variants/libfw_*.so; do
  r=$(FW_B200_LIB=$PWD/$f python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1)
  echo "$f $(echo "$r" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("update_ms=%.4f frac=%.3f ms_per_step=%.4f plan=%.4f spawn=%.4f"%(d["kernel_ms"]["update"], d["roofline"]["frac"], d["ms_per_step"], d["kernel_ms"]["plan"], d["kernel_ms"]["spawn"]))' 2>&1 | tail -1)"
done
