#!/bin/bash
# round 2, call 37: length hint + tile-ordered kernel as dispatched: lev tests incl. forced variants, bench lines (headline must be unchanged)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev or test_length_hint or test_full_size_properties or test_nul or bitpar-duo or test_abi or dev" 2>&1 | tail -4 > gpurun_out/r02_c37_tests.txt
cat gpurun_out/r02_c37_tests.txt
rm -f gpurun_out/r02_c37_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c37_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 100 --warmup 5 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'], 'e2e_ms', d['e2e']['ms_per_step'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c37_bench.txt
}
run lev_k8_len128 TA_X=1
run lev_k8_ragged96_160 TA_X=1
run lev_k8_len128_R TA_X=1
cat gpurun_out/r02_c37_bench.txt
