#!/bin/bash
# round 2, first GPU batch: parity suite with the shared sincos + full-size tests, baseline bench
mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpu.txt 2>&1
nproc >> gpurun_out/r2a/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r2a/pytest_gpu.log 2>&1
tail -30 gpurun_out/r2a/pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/r2a/bench_c3.json 2> gpurun_out/r2a/bench_c3.err
tail -c 1500 gpurun_out/r2a/bench_c3.json
