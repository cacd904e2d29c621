#!/bin/bash
# round 2, call 40: headline kernel with 2 resident CTAs per SM and a smaller shared-memory carve-out (more L1)
mkdir -p gpurun_out; rm -f gpurun_out/r02_c40_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c40_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 200 --warmup 5 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c40_bench.txt
}
run lev_k8_len128 TA_X=1
run lev_k8_len128 TA_DUO_CTAS=2
run lev_k8_len128 TA_DUO_CTAS=2 TA_DUO_CARVEOUT=58
run lev_k8_len128 TA_DUO_CTAS=2 TA_DUO_CARVEOUT=72
run lev_k8_len128 TA_DUO_CARVEOUT=100
run lev_k8_len128 TA_DUO_CARVEOUT=86
cat gpurun_out/r02_c40_bench.txt
