#!/bin/bash
# round 2, call 41: compute-sanitizer over the round-2 kernels incl. the tile-ordered two-pairs-per-thread kernel
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_round2.py > gpurun_out/r02_c41_memcheck.log 2>&1; tail -3 gpurun_out/r02_c41_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize_round2.py > gpurun_out/r02_c41_racecheck.log 2>&1; tail -3 gpurun_out/r02_c41_racecheck.log
timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize_round2.py > gpurun_out/r02_c41_synccheck.log 2>&1; tail -3 gpurun_out/r02_c41_synccheck.log
