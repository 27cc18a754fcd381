#!/bin/bash
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
FW_FUZZ_SEEDS=5-40 timeout 1200 python -m pytest tests/test_gpu_fuzz_nested.py -m gpu -q > $O/fuzz_nested.log 2>&1; tail -40 $O/fuzz_nested.log
