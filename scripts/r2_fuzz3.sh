#!/bin/bash
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
FW_FUZZ_SEEDS=41-95 timeout 480 python -m pytest tests/test_gpu_fuzz_nested.py -m gpu -q > $O/fuzz_nested_final.log 2>&1; tail -4 $O/fuzz_nested_final.log
FW_FUZZ_SEEDS=161-240 timeout 400 python -m pytest tests/test_gpu_edge_and_scale.py -m gpu -q -k randomized_mixed_scene > $O/fuzz_mixed_final.log 2>&1; tail -4 $O/fuzz_mixed_final.log
