#!/bin/bash
mkdir -p gpurun_out/r2x3; O=gpurun_out/r2x3
K="cylinder_and_cone or many_slow_emitters or nested_emission_textures or collision_destroy or random_lifetime_compaction or sparks_trajectory"
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -x -q -k "cylinder_and_cone" 2>&1 | tail -6) > $O/memcheck_capsule.log; cat $O/memcheck_capsule.log
(timeout 900 compute-sanitizer --tool synccheck python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -6) > $O/synccheck.log; cat $O/synccheck.log
(timeout 900 compute-sanitizer --tool initcheck python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -25) > $O/initcheck.log; cat $O/initcheck.log
