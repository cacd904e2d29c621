#!/bin/bash
# round 2, call 13: compute-sanitizer over the round-2 kernels; occupancy variants of the diagonal-extension kernel
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_round2.py > gpurun_out/r02_c13_memcheck.log 2>&1; tail -4 gpurun_out/r02_c13_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_round2.py > gpurun_out/r02_c13_racecheck.log 2>&1; tail -4 gpurun_out/r02_c13_racecheck.log
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c13_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c13_bench.txt
}
for occ in 8 9 10 12; do
run lev_k16_len4096 TA_FR_OCC=$occ
run exp_len1024 TA_FR_OCC=$occ
done
cat gpurun_out/r02_c13_bench.txt
