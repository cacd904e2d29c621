#!/bin/bash
# compute-sanitizer over the kernels added in the second half of the round
mkdir -p gpurun_out
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py ) > gpurun_out/c16_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/c16_memcheck.log
( time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py ) > gpurun_out/c16_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/c16_racecheck.log
TA_PIGEON_STAGED=0 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/c16_memcheck_unstaged.log 2>&1
echo "memcheck (lane-per-segment filter) rc=$?"; tail -2 gpurun_out/c16_memcheck_unstaged.log
