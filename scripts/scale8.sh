#!/bin/bash
# 8-GPU box: weak + strong scaling of bench.py at N=8, then the render-extract gather at N=8
mkdir -p gpurun_out/final
bash scripts/scale_run.sh 8 > /dev/null 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/gather_bench.py c3 2>&1 | tail -1 > gpurun_out/final/gather_8.json
cp gpurun_out/s2_scale_weak_8.json gpurun_out/final/scale_weak_8.json; cp gpurun_out/s2_scale_strong_8.json gpurun_out/final/scale_strong_8.json
tail -c 400 gpurun_out/final/gather_8.json
