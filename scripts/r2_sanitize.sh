#!/bin/bash
mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "layout or handoff or compaction or nested or destroyed or edge or (randomized_mixed_scene and 1) or sincos or collision_single or c2_1m" 2>&1 | tail -8) > $O/memcheck.log; cat $O/memcheck.log
(timeout 900 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -x -q -k "layout or compaction or nested or destroyed or collision_destroy or (randomized_mixed_scene and 1)" 2>&1 | tail -8) > $O/racecheck.log; cat $O/racecheck.log
timeout 600 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -2
