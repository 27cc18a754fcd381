#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_edge_and_scale.py -m gpu -q -k "profile_counters or long_run" 2>&1 | tail -30
