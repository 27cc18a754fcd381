#!/bin/bash
mkdir -p gpurun_out/r2u; O=gpurun_out/r2u
timeout 300 python scripts/host_cost_probe.py > $O/host_cost.txt 2>&1; cat $O/host_cost.txt
