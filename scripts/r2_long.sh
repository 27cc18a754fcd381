#!/bin/bash
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
FW_LONG_FRAMES=20000 timeout 1200 python -m pytest tests/test_gpu_edge_and_scale.py -m gpu -q -k long_run > $O/long_run.log 2>&1; tail -30 $O/long_run.log
timeout 600 python -m pytest tests/test_gpu_edge_and_scale.py -m gpu -q -k long_run --durations=3 2>&1 | tail -8
