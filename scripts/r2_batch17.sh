#!/bin/bash
mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
FW_B200_LIB=build_variants/libfw_stats.so timeout 300 python scripts/collide_stats.py c5 > $O/stats_c5.txt 2>&1; cat $O/stats_c5.txt
FW_B200_LIB=build_variants/libfw_stats.so timeout 300 python scripts/collide_stats.py c5d > $O/stats_c5d.txt 2>&1; cat $O/stats_c5d.txt
