#!/bin/bash
# round 2, call 29: two-pairs-per-thread kernel behind a tile-local ordering by length class: parity and timing vs the old kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_k_mutated or test_lev_k_random_short or test_nul_bytes or test_lev_exp or test_lev_duo_ragged_tiles or test_full_size_properties or bitpar-duo or length-bucketing" 2>&1 | tail -4 > gpurun_out/r02_c29_tests.txt
cat gpurun_out/r02_c29_tests.txt
rm -f gpurun_out/r02_c29_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c29_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 200 --warmup 5 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c29_bench.txt
}
for rep in 1 2; do
run lev_k8_len128 TA_DUO_TILED=1
run lev_k8_len128 TA_DUO_TILED=0
done
run lev_k8_ragged96_160 TA_DUO_TILED=1
run lev_k8_ragged96_160 TA_DUO_TILED=0
run lev_k8_ragged96_160 TA_DUO_TILED=0 TA_LEN_BUCKETS=1
run lev_k8_len128_R TA_DUO_TILED=1
run lev_k8_len128_R TA_DUO_TILED=0
cat gpurun_out/r02_c29_bench.txt
