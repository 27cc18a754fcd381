#!/bin/bash
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
FW_FUZZ_SEEDS=7-160 timeout 1500 python -m pytest tests/test_gpu_edge_and_scale.py -m gpu -q -k randomized_mixed_scene > $O/fuzz.log 2>&1; tail -15 $O/fuzz.log
