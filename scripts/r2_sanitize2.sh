#!/bin/bash
mkdir -p gpurun_out/r2x; O=gpurun_out/r2x
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "many_slow_emitters or zero_dt or large_one_shot or nested or sparks_trajectory or collision_destroy" 2>&1 | tail -8) > $O/memcheck.log; cat $O/memcheck.log
(timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "many_slow_emitters or zero_dt or nested or sparks_trajectory" 2>&1 | tail -8) > $O/racecheck.log; cat $O/racecheck.log
